// sp_conv_gemm.cuh — GEMM tier for the wide, spatially tiny bottleneck layers of the auto-encoder: Conv3d 32->100 k3 s2
// (7x25x25 -> 3x12x12), Conv3d 100->fc k3 (3x12x12 -> 1x10x10), ConvTranspose3d fc->100 k3 and 100->32 k3 s2
// (Cae3D.py:70,74,178,182).  Their planes are too small for the spatial tiles of sp_conv_tiled.cuh, but with K = 27*Ci =
// 864..5400 and Co = 32..800 they are plain dense contractions:
//   sp_corr  : O[(n,o)][co]      = A[(n,o)][(tap,ci)] * Wc[(tap,ci)][co]           A = im2col(BN(I)), zero padding
//   sp_corrT : P[(n,o)][(tap,ci)] = O'[(n,o)][co] * Wt[co][(tap,ci)]  then  I[n,i,ci] = sum_tap P[(n,(i+p-tap)/s)][tap][ci]
//   sp_wgrad : dW'[(tap,ci)][co] = sum_(n,o) A[(n,o)][(tap,ci)] * O'[(n,o)][co]     (split over the rows, fixed-order reduce)
// The matrices live in a caller-provided workspace (sp_conv_workspace_bytes / sp_wgrad_workspace_bytes).  The contraction
// is a 64x64x16 shared-memory tiled fp32 FFMA GEMM (4x4 register tile per thread): exact fp32, deterministic.
#pragma once
#include "sp_common.cuh"

namespace sp_gemm {

constexpr int BM = 64, BN = 64, BK = 16;

// C[M][N] (ldc) = A * B over the reduction range of blockIdx.z.  TA = 0: A is [M][K] (lda); TA = 1: A is [K][M] (lda).
// B is [K][N] (ldb).  EPI = 1: C = act(C + bias[n]);  EPI = 0: plain partial, slice z written at C + z * slice_stride.
template <int TA, int EPI>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, int k_per_slice, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
             float* __restrict__ C, int ldc, int64_t slice_stride, const float* __restrict__ bias, int act, float alpha) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_slice;
    const int kend = (kbeg + k_per_slice < K) ? kbeg + k_per_slice : K;
    const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool va = (lda % 4 == 0), vb = (ldb % 4 == 0);

    // Register double buffering: the global loads of k-tile i + 1 are issued before the FFMAs of k-tile i (the kernel runs with
    // a handful of warps per SM on the bottleneck layers: without the prefetch every k-tile paid a full load latency).
    float ra[4];
    float4 rb;
    auto load_tiles = [&](int k0) {
        if (TA == 0) {
            const int row = t >> 2, kq = (t & 3) * 4;
            const int m = m0 + row, k = k0 + kq;
            ra[0] = ra[1] = ra[2] = ra[3] = 0.f;
            if (m < M) {
                const float* p = A + (int64_t)m * lda + k;
                if (va && k + 3 < kend) {
                    const float4 q = *reinterpret_cast<const float4*>(p);
                    ra[0] = q.x; ra[1] = q.y; ra[2] = q.z; ra[3] = q.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (k + u < kend) ra[u] = p[u];
                }
            }
        } else {
            const int krow = t >> 4, mq = (t & 15) * 4;
            const int k = k0 + krow, m = m0 + mq;
            ra[0] = ra[1] = ra[2] = ra[3] = 0.f;
            if (k < kend) {
                const float* p = A + (int64_t)k * lda + m;
                if (va && m + 3 < M) {
                    const float4 q = *reinterpret_cast<const float4*>(p);
                    ra[0] = q.x; ra[1] = q.y; ra[2] = q.z; ra[3] = q.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (m + u < M) ra[u] = p[u];
                }
            }
        }
        {
            const int krow = t >> 4, nq = (t & 15) * 4;
            const int k = k0 + krow, n = n0 + nq;
            rb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < kend) {
                const float* p = B + (int64_t)k * ldb + n;
                if (vb && n + 3 < N) {
                    rb = *reinterpret_cast<const float4*>(p);
                } else {
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (n + u < N) v[u] = p[u];
                    rb = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    };
    auto store_tiles = [&]() {
        if (TA == 0) {
            const int row = t >> 2, kq = (t & 3) * 4;
#pragma unroll
            for (int u = 0; u < 4; ++u) As[kq + u][row] = ra[u];
        } else {
            const int krow = t >> 4, mq = (t & 15) * 4;
            *reinterpret_cast<float4*>(&As[krow][mq]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
        }
        const int krow = t >> 4, nq = (t & 15) * 4;
        *reinterpret_cast<float4*>(&Bs[krow][nq]) = rb;
    };

    if (kbeg < kend) load_tiles(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < kend) load_tiles(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* Cz = C + (int64_t)blockIdx.z * slice_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (EPI == 1) v = sp_act_fwd(v + (bias ? bias[n] : 0.f), act, alpha);
            Cz[(int64_t)m * ldc + n] = v;
        }
    }
}

// A[(n,o)][tap*Ci + ci] = BN(I[n, o*s - p + tap, ci]) (0 outside the volume).  One thread per (row, tap, channel).
__global__ void __launch_bounds__(256)
im2col_kernel(SpConvDesc d, int nPerG, const float* __restrict__ src, const float* __restrict__ scale, const float* __restrict__ shift,
              float* __restrict__ A) {
    const int k3 = d.k * d.k * d.k;
    const int64_t total = (int64_t)d.N * d.Do * d.Ho * d.Wo * k3 * d.Ci;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % d.Ci);
        int64_t r = i / d.Ci;
        const int tap = (int)(r % k3); r /= k3;
        const int ow = (int)(r % d.Wo); r /= d.Wo;
        const int oh = (int)(r % d.Ho); r /= d.Ho;
        const int od = (int)(r % d.Do);
        const int n = (int)(r / d.Do);
        const int kw = tap % d.k, kh = (tap / d.k) % d.k, kd = tap / (d.k * d.k);
        const int id = od * d.s - d.pd + kd, ih = oh * d.s - d.ph + kh, iw = ow * d.s - d.pw + kw;
        float v = 0.f;
        if (id >= 0 && id < d.Di && ih >= 0 && ih < d.Hi && iw >= 0 && iw < d.Wi) {
            v = src[((((int64_t)n * d.Di + id) * d.Hi + ih) * d.Wi + iw) * d.ldi + ci];
            if (scale) {
                const int g = n / nPerG;
                v = fmaf(v, scale[(int64_t)g * d.Ci + ci], shift[(int64_t)g * d.Ci + ci]);
            }
        }
        A[i] = v;
    }
}

// The same matrix, one thread per (row, tap, channel QUAD): 128-bit loads / stores and 32-bit index arithmetic (the scalar form
// spends its time in six 64-bit divisions per element: 100 us for the 3200 x 2700 matrix of Cae3D.py:74 at batch 32).
// Requires Ci % 4 == 0, ldi % 4 == 0, 16-byte aligned pointers and fewer than 2^31 quads.
__global__ void __launch_bounds__(256)
im2col_quad_kernel(SpConvDesc d, int nPerG, const float* __restrict__ src, const float* __restrict__ scale, const float* __restrict__ shift,
                   float* __restrict__ A) {
    const int k3 = d.k * d.k * d.k, cq = d.Ci >> 2;
    const int total = d.N * d.Do * d.Ho * d.Wo * k3 * cq;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int q = i % cq;
        int r = i / cq;
        const int tap = r % k3; r /= k3;
        const int ow = r % d.Wo; r /= d.Wo;
        const int oh = r % d.Ho; r /= d.Ho;
        const int od = r % d.Do;
        const int n = r / d.Do;
        const int kw = tap % d.k, kh = (tap / d.k) % d.k, kd = tap / (d.k * d.k);
        const int id = od * d.s - d.pd + kd, ih = oh * d.s - d.ph + kh, iw = ow * d.s - d.pw + kw;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id >= 0 && id < d.Di && ih >= 0 && ih < d.Hi && iw >= 0 && iw < d.Wi) {
            v = *reinterpret_cast<const float4*>(src + ((((int64_t)n * d.Di + id) * d.Hi + ih) * d.Wi + iw) * d.ldi + 4 * q);
            if (scale) {
                const int g = n / nPerG;
                const float4 sc = *reinterpret_cast<const float4*>(scale + (int64_t)g * d.Ci + 4 * q);
                const float4 sh = *reinterpret_cast<const float4*>(shift + (int64_t)g * d.Ci + 4 * q);
                v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
            }
        }
        reinterpret_cast<float4*>(A)[i] = v;
    }
}

static inline bool im2col_quad_ok(const SpConvDesc* d, const void* src, const void* scale, const void* shift, const void* A) {
    const int64_t quads = (int64_t)d->N * d->Do * d->Ho * d->Wo * d->k * d->k * d->k * (d->Ci / 4);
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift) |
                         reinterpret_cast<uintptr_t>(A);
    return d->Ci % 4 == 0 && d->ldi % 4 == 0 && (al & 15) == 0 && quads < (1LL << 31) - (1 << 22);
}

// dense BN-applied copy of the O-side: Oc[(n,o)][co]
__global__ void __launch_bounds__(256)
oside_copy_kernel(SpConvDesc d, int nPerG, const float* __restrict__ src, const float* __restrict__ scale,
                  const float* __restrict__ shift, float* __restrict__ Oc) {
    const int64_t vox = (int64_t)d.Do * d.Ho * d.Wo;
    const int64_t total = (int64_t)d.N * vox * d.Co;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % d.Co);
        const int64_t m = i / d.Co;
        float v = src[m * d.ldo + co];
        if (scale) {
            const int g = (int)(m / vox) / nPerG;
            v = fmaf(v, scale[(int64_t)g * d.Co + co], shift[(int64_t)g * d.Co + co]);
        }
        Oc[i] = v;
    }
}

// I[n,i,ci] = act(bias[ci] + sum over taps with (i + p - tap) % s == 0 of P[(n,(i+p-tap)/s)][tap][ci]); P row stride ldp
__global__ void __launch_bounds__(256)
col2im_kernel(SpConvDesc d, const float* __restrict__ P, int ldp, int ciP, const float* __restrict__ bias, float* __restrict__ dst) {
    const int64_t total = (int64_t)d.N * d.Di * d.Hi * d.Wi * d.Ci;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % d.Ci);
        int64_t r = i / d.Ci;
        const int iw = (int)(r % d.Wi); r /= d.Wi;
        const int ih = (int)(r % d.Hi); r /= d.Hi;
        const int id = (int)(r % d.Di);
        const int n = (int)(r / d.Di);
        float acc = 0.f;
        for (int kd = 0; kd < d.k; ++kd) {
            const int td = id + d.pd - kd;
            if (td < 0 || (td % d.s) != 0 || td / d.s >= d.Do) continue;
            for (int kh = 0; kh < d.k; ++kh) {
                const int th = ih + d.ph - kh;
                if (th < 0 || (th % d.s) != 0 || th / d.s >= d.Ho) continue;
                for (int kw = 0; kw < d.k; ++kw) {
                    const int tw = iw + d.pw - kw;
                    if (tw < 0 || (tw % d.s) != 0 || tw / d.s >= d.Wo) continue;
                    const int tap = (kd * d.k + kh) * d.k + kw;
                    const int64_t m = (((int64_t)n * d.Do + td / d.s) * d.Ho + th / d.s) * d.Wo + tw / d.s;
                    acc += P[m * ldp + tap * ciP + ci];
                }
            }
        }
        dst[(i / d.Ci) * d.ldi + ci] = sp_act_fwd(acc + (bias ? bias[ci] : 0.f), d.act, d.alpha);
    }
}

// The same gather, one thread per (voxel, channel QUAD), 32-bit index arithmetic, the tap loops without divisions for stride 1 / 2.
// Requires Ci % 4 == 0, ldi % 4 == 0, ciP % 4 == 0, ldp % 4 == 0, 16-byte aligned pointers, fewer than 2^31 quads.
__global__ void __launch_bounds__(256)
col2im_quad_kernel(SpConvDesc d, const float* __restrict__ P, int ldp, int ciP, const float* __restrict__ bias, float* __restrict__ dst) {
    const int cq = d.Ci >> 2;
    const int total = d.N * d.Di * d.Hi * d.Wi * cq;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int q = i % cq;
        int r = i / cq;
        const int vox = r;
        const int iw = r % d.Wi; r /= d.Wi;
        const int ih = r % d.Hi; r /= d.Hi;
        const int id = r % d.Di;
        const int n = r / d.Di;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kd = 0; kd < d.k; ++kd) {
            const int td = id + d.pd - kd;
            if (td < 0 || (d.s == 2 && (td & 1)) || td / d.s >= d.Do) continue;
            for (int kh = 0; kh < d.k; ++kh) {
                const int th = ih + d.ph - kh;
                if (th < 0 || (d.s == 2 && (th & 1)) || th / d.s >= d.Ho) continue;
                for (int kw = 0; kw < d.k; ++kw) {
                    const int tw = iw + d.pw - kw;
                    if (tw < 0 || (d.s == 2 && (tw & 1)) || tw / d.s >= d.Wo) continue;
                    const int tap = (kd * d.k + kh) * d.k + kw;
                    const int64_t m = (((int64_t)n * d.Do + td / d.s) * d.Ho + th / d.s) * d.Wo + tw / d.s;
                    const float4 p = *reinterpret_cast<const float4*>(P + m * ldp + tap * ciP + 4 * q);
                    acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
                }
            }
        }
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b = *reinterpret_cast<const float4*>(bias + 4 * q);
        float4 o;
        o.x = sp_act_fwd(acc.x + b.x, d.act, d.alpha); o.y = sp_act_fwd(acc.y + b.y, d.act, d.alpha);
        o.z = sp_act_fwd(acc.z + b.z, d.act, d.alpha); o.w = sp_act_fwd(acc.w + b.w, d.act, d.alpha);
        *reinterpret_cast<float4*>(dst + (int64_t)vox * d.ldi + 4 * q) = o;
    }
}

static inline bool col2im_quad_ok(const SpConvDesc* d, const void* P, int ldp, int ciP, const void* bias, const void* dst) {
    const int64_t quads = (int64_t)d->N * d->Di * d->Hi * d->Wi * (d->Ci / 4);
    const uintptr_t al = reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(dst);
    return d->Ci % 4 == 0 && d->ldi % 4 == 0 && ldp % 4 == 0 && ciP % 4 == 0 && (d->s == 1 || d->s == 2) && (al & 15) == 0 &&
           quads < (1LL << 31) - (1 << 22);
}

// dw[co][ci][tap] = beta*dw + sum_z part[z][(tap,ci)][co]
__global__ void __launch_bounds__(256)
wgrad_fold_kernel(const float* __restrict__ part, int slices, int k3, int Ci, int Co, float* __restrict__ dw, float beta) {
    const int64_t wn = (int64_t)Co * Ci * k3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < wn; i += (int64_t)gridDim.x * blockDim.x) {
        const int tap = (int)(i % k3);
        const int ci = (int)((i / k3) % Ci);
        const int co = (int)(i / ((int64_t)k3 * Ci));
        double s = 0.0;
        for (int z = 0; z < slices; ++z) s += (double)part[((int64_t)z * k3 * Ci + (int64_t)tap * Ci + ci) * Co + co];
        dw[i] = (beta == 0.f) ? (float)s : fmaf(beta, dw[i], (float)s);
    }
}

// dst[m][n] = act(bias[n] + sum_z part[z][m][n]): epilogue of the split-K forward (fixed order over the slices)
__global__ void __launch_bounds__(256)
corr_fold_kernel(const float* __restrict__ part, int slices, int64_t M, int N, const float* __restrict__ bias, int act, float alpha,
                 float* __restrict__ dst, int ldc) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M * N; i += (int64_t)gridDim.x * blockDim.x) {   // ew_blocks caps the grid
        const int64_t m = i / N;
        const int n = (int)(i - m * N);
        float v = 0.f;
        for (int z = 0; z < slices; ++z) v += part[(int64_t)z * M * N + i];
        dst[m * ldc + n] = sp_act_fwd(v + (bias ? bias[n] : 0.f), act, alpha);
    }
}

// which = 1 pack of the GEMM tier: Wt[co][tap][ciP]
__global__ void pack_wt_gemm_kernel(const float* __restrict__ w, float* __restrict__ wp, int Co, int Ci, int k3, int ciP) {
    const int64_t total = (int64_t)Co * k3 * ciP;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % ciP);
        const int tap = (int)((i / ciP) % k3);
        const int co = (int)(i / ((int64_t)ciP * k3));
        wp[i] = (ci < Ci) ? w[((int64_t)co * Ci + ci) * k3 + tap] : 0.f;
    }
}

static inline int ew_blocks(int64_t total) {
    int64_t b = (total + 255) / 256;
    const int64_t cap = (int64_t)sp_num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

struct Plan {
    int64_t M;          // rows (n, o)
    int kdim, k3, ciP;  // tap*Ci + ci columns; taps; padded I-side channels of the corrT pack
    int slices;         // split of the rows for wgrad
    int64_t rows_per_slice;
    int fslices, fk_per_slice;   // split of K = kdim for the forward (the tiles alone are 1..2 CTAs per SM: latency-bound)
};

static inline Plan plan(const SpConvDesc* d) {
    Plan p;
    p.k3 = d->k * d->k * d->k;
    p.M = (int64_t)d->N * d->Do * d->Ho * d->Wo;
    p.kdim = p.k3 * d->Ci;
    p.ciP = (d->Ci + 15) / 16 * 16;
    const int64_t tiles = sp_cdiv(p.kdim, BM) * sp_cdiv(d->Co, BN);
    int64_t s = sp_cdiv(2 * (int64_t)sp_num_sms(), tiles);
    const int64_t smax = sp_cdiv(p.M, 64);
    if (s > smax) s = smax;
    if (s > 32) s = 32;
    if (s < 1) s = 1;
    p.rows_per_slice = sp_cdiv(sp_cdiv(p.M, s), BK) * BK;
    p.slices = (int)sp_cdiv(p.M, p.rows_per_slice);
    {
        const int64_t ftiles = sp_cdiv(p.M, BM) * sp_cdiv(d->Co, BN);
        int64_t fs = sp_cdiv(4 * (int64_t)sp_num_sms(), ftiles);
        static int nosplit = -1;   // SP_GEMM_NOSPLITK=1: single-slice forward (A/B checks)
        if (nosplit < 0) {
            const char* e = getenv("SP_GEMM_NOSPLITK");
            nosplit = (e && e[0] == '1') ? 1 : 0;
        }
        if (nosplit) fs = 1;
        if (fs > p.kdim / 128) fs = p.kdim / 128;
        if (fs > 16) fs = 16;
        if (fs < 1) fs = 1;
        p.fk_per_slice = (int)(sp_cdiv(sp_cdiv(p.kdim, fs), BK) * BK);
        p.fslices = (int)sp_cdiv(p.kdim, p.fk_per_slice);
    }
    return p;
}

}  // namespace sp_gemm

static inline bool sp_gemm_disabled() {
    static int v = -1;   // SP_DISABLE_GEMM=1 forces the generic kernels
    if (v < 0) {
        const char* e = getenv("SP_DISABLE_GEMM");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// wide, spatially small layers that neither spatial tier serves
static inline bool sp_gemm_serves(const SpConvDesc* d) {
    if (sp_gemm_disabled() || d->k < 2) return false;
    if (sp_tiled_corr_supported(d)) return false;
    return d->Ci * d->k * d->k * d->k >= 256 && d->Co >= 16 && d->Ci >= 8;
}

static inline size_t sp_gemm_corr_ws_bytes(const SpConvDesc* d) {
    const sp_gemm::Plan p = sp_gemm::plan(d);
    return ((size_t)p.M * p.kdim + (p.fslices > 1 ? (size_t)p.fslices * p.M * d->Co : 0)) * sizeof(float);
}
static inline size_t sp_gemm_corrT_ws_bytes(const SpConvDesc* d) {
    const sp_gemm::Plan p = sp_gemm::plan(d);
    return ((size_t)p.M * d->Co + (size_t)p.M * p.k3 * p.ciP) * sizeof(float);
}
static inline size_t sp_gemm_wgrad_ws_bytes(const SpConvDesc* d) {
    const sp_gemm::Plan p = sp_gemm::plan(d);
    return ((size_t)p.M * p.kdim + (size_t)p.M * d->Co + (size_t)p.slices * p.kdim * d->Co) * sizeof(float);
}

static inline int sp_gemm_corr_launch(const SpConvDesc* d, int nPerG, const float* src, const float* wp, const float* bias,
                                      const float* scale, const float* shift, float* dst, float* ws, cudaStream_t st) {
    using namespace sp_gemm;
    const Plan p = plan(d);
    const int coP = (d->Co + 15) / 16 * 16;
    if (im2col_quad_ok(d, src, scale, shift, ws)) im2col_quad_kernel<<<ew_blocks(p.M * p.kdim / 4), 256, 0, st>>>(*d, nPerG, src, scale, shift, ws);
    else im2col_kernel<<<ew_blocks(p.M * p.kdim), 256, 0, st>>>(*d, nPerG, src, scale, shift, ws);
    SP_LAUNCH_OK("im2col_kernel");
    if (p.fslices <= 1) {
        dim3 grid((unsigned)sp_cdiv(d->Co, BN), (unsigned)sp_cdiv(p.M, BM), 1);
        sgemm_kernel<0, 1><<<grid, 256, 0, st>>>((int)p.M, d->Co, p.kdim, p.kdim, ws, p.kdim, wp, coP, dst, d->ldo, 0, bias, d->act, d->alpha);
        SP_LAUNCH_OK("sgemm_kernel");
        return 0;
    }
    // split K: partial products of every K slice, then bias + activation in the fold (fixed order)
    float* part = ws + (size_t)p.M * p.kdim;
    dim3 grid((unsigned)sp_cdiv(d->Co, BN), (unsigned)sp_cdiv(p.M, BM), (unsigned)p.fslices);
    sgemm_kernel<0, 0><<<grid, 256, 0, st>>>((int)p.M, d->Co, p.kdim, p.fk_per_slice, ws, p.kdim, wp, coP, part, d->Co, (int64_t)p.M * d->Co, nullptr, 0, 0.f);
    SP_LAUNCH_OK("sgemm_kernel");
    corr_fold_kernel<<<ew_blocks(p.M * d->Co), 256, 0, st>>>(part, p.fslices, p.M, d->Co, bias, d->act, d->alpha, dst, d->ldo);
    SP_LAUNCH_OK("corr_fold_kernel");
    return 0;
}

static inline int sp_gemm_corrT_launch(const SpConvDesc* d, int nPerG, const float* src, const float* wp, const float* bias,
                                       const float* scale, const float* shift, float* dst, float* ws, cudaStream_t st) {
    using namespace sp_gemm;
    const Plan p = plan(d);
    float* Oc = ws;
    float* P = ws + (size_t)p.M * d->Co;
    const int ncol = p.k3 * p.ciP;
    oside_copy_kernel<<<ew_blocks(p.M * d->Co), 256, 0, st>>>(*d, nPerG, src, scale, shift, Oc);
    SP_LAUNCH_OK("oside_copy_kernel");
    dim3 grid((unsigned)sp_cdiv(ncol, BN), (unsigned)sp_cdiv(p.M, BM), 1);
    sgemm_kernel<0, 0><<<grid, 256, 0, st>>>((int)p.M, ncol, d->Co, d->Co, Oc, d->Co, wp, ncol, P, ncol, 0, nullptr, 0, 0.f);
    SP_LAUNCH_OK("sgemm_kernel");
    if (col2im_quad_ok(d, P, ncol, p.ciP, bias, dst))
        col2im_quad_kernel<<<ew_blocks((int64_t)d->N * d->Di * d->Hi * d->Wi * d->Ci / 4), 256, 0, st>>>(*d, P, ncol, p.ciP, bias, dst);
    else
        col2im_kernel<<<ew_blocks((int64_t)d->N * d->Di * d->Hi * d->Wi * d->Ci), 256, 0, st>>>(*d, P, ncol, p.ciP, bias, dst);
    SP_LAUNCH_OK("col2im_kernel");
    return 0;
}

static inline int sp_gemm_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                       const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                       cudaStream_t st) {
    using namespace sp_gemm;
    const Plan p = plan(d);
    float* A = ws;
    float* Oc = A + (size_t)p.M * p.kdim;
    float* part = Oc + (size_t)p.M * d->Co;
    if (im2col_quad_ok(d, iside, i_scale, i_shift, A)) im2col_quad_kernel<<<ew_blocks(p.M * p.kdim / 4), 256, 0, st>>>(*d, nPerG, iside, i_scale, i_shift, A);
    else im2col_kernel<<<ew_blocks(p.M * p.kdim), 256, 0, st>>>(*d, nPerG, iside, i_scale, i_shift, A);
    SP_LAUNCH_OK("im2col_kernel");
    oside_copy_kernel<<<ew_blocks(p.M * d->Co), 256, 0, st>>>(*d, nPerG, oside, o_scale, o_shift, Oc);
    SP_LAUNCH_OK("oside_copy_kernel");
    // part[z][(tap,ci)][co] = sum over the rows of slice z of A[m][(tap,ci)] * Oc[m][co]   (A^T * Oc, reduction = rows)
    dim3 grid((unsigned)sp_cdiv(d->Co, BN), (unsigned)sp_cdiv(p.kdim, BM), (unsigned)p.slices);
    sgemm_kernel<1, 0><<<grid, 256, 0, st>>>(p.kdim, d->Co, (int)p.M, (int)p.rows_per_slice, A, p.kdim, Oc, d->Co, part, d->Co,
                                             (int64_t)p.kdim * d->Co, nullptr, 0, 0.f);
    SP_LAUNCH_OK("sgemm_kernel");
    wgrad_fold_kernel<<<ew_blocks((int64_t)d->Co * p.kdim), 256, 0, st>>>(part, p.slices, p.k3, d->Ci, d->Co, dw, beta);
    SP_LAUNCH_OK("wgrad_fold_kernel");
    return 0;
}
