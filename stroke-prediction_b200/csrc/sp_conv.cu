// sp_conv.cu — correlation / transposed correlation / weight-gradient kernels (fp32, NDHWC) and their C-ABI.
//
// Reference call sites replaced: nn.Conv3d and nn.ConvTranspose3d forward/backward at
// common/model/Cae3D.py:41-74,126-132,178-218 and common/model/Unet3D.py:19,22,50,52 (see include/stroke_b200.h).
//
// Two kernel tiers:
//   * generic kernels (this file, *_generic_kernel): any k in {1,2,3}, s in {1,2}, any padding / channel count.
//     One thread owns one destination voxel x 8 destination channels and walks the taps straight from global
//     memory.  They are the correctness baseline and serve the HBM-bound layers (1x1x1, k2s2, C<=3).
//   * tiled kernels (sp_conv_tiled.cuh): 3x3x3 stride-1 correlation on shared-memory halo tiles with an 8x8
//     register micro-tile per thread — the FFMA-bound 16..96-channel layers.
#include "sp_common.cuh"
#include "sp_conv_tiled.cuh"
#include "sp_conv_tiledT.cuh"
#include "sp_conv_pw.cuh"
#include "sp_conv_gemm.cuh"
#include "sp_conv_tc.cuh"
#include "sp_conv_tc2.cuh"
#include "sp_conv_tc3.cuh"
#include "sp_wgrad_tc.cuh"
#include "sp_wgrad_tc24.cuh"
#include "sp_wgrad_tc4.cuh"
#include "sp_wgrad_tc4s2.cuh"
#include "sp_conv_thin.cuh"
#include "sp_conv_k2s2.cuh"

// fixed-order sum of per-CTA partial weight-gradient slabs (fp64 accumulation: the partials carry the rounding of long
// fp32 chains already, the cross-CTA sum should not add to it)
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int chunks, int64_t wn, float* __restrict__ dw, float beta) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < wn; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < chunks; ++c) s += (double)ws[(int64_t)c * wn + i];
        dw[i] = (beta == 0.f) ? (float)s : fmaf(beta, dw[i], (float)s);
    }
}

namespace {

constexpr int kPad = 16;  // packed weights pad the fastest (destination-channel) axis to a multiple of 16

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

int check_desc(const SpConvDesc* d) {
    SP_REQUIRE(d != nullptr, "SpConvDesc is NULL");
    SP_REQUIRE(d->N > 0 && d->Ci > 0 && d->Co > 0, "conv: N, Ci, Co must be positive (N=%d Ci=%d Co=%d)", d->N, d->Ci, d->Co);
    SP_REQUIRE(d->k >= 1 && d->k <= 3, "conv: kernel extent %d not in {1,2,3}", d->k);
    SP_REQUIRE(d->s == 1 || d->s == 2, "conv: stride %d not in {1,2}", d->s);
    SP_REQUIRE(d->pd >= 0 && d->ph >= 0 && d->pw >= 0, "conv: negative padding");
    SP_REQUIRE(d->ldi >= d->Ci && d->ldo >= d->Co, "conv: ld smaller than channel count");
    SP_REQUIRE(d->Di > 0 && d->Hi > 0 && d->Wi > 0 && d->Do > 0 && d->Ho > 0 && d->Wo > 0, "conv: empty extent");
    // O-side extents are the floor-mode output size of the I-side (the same relation ConvTranspose3d inverts).
    SP_REQUIRE(d->Di + 2 * d->pd >= d->k && d->Hi + 2 * d->ph >= d->k && d->Wi + 2 * d->pw >= d->k,
               "conv: kernel larger than padded I-side");
    SP_REQUIRE(d->Do == (d->Di + 2 * d->pd - d->k) / d->s + 1 && d->Ho == (d->Hi + 2 * d->ph - d->k) / d->s + 1 &&
                   d->Wo == (d->Wi + 2 * d->pw - d->k) / d->s + 1,
               "conv: O-side %dx%dx%d inconsistent with I-side %dx%dx%d (k=%d s=%d p=%d,%d,%d)", d->Do, d->Ho, d->Wo,
               d->Di, d->Hi, d->Wi, d->k, d->s, d->pd, d->ph, d->pw);
    SP_REQUIRE(d->act >= SP_ACT_NONE && d->act <= SP_ACT_SIGMOID, "conv: unknown activation %d", d->act);
    return 0;
}

// ---- weight packing -----------------------------------------------------------------------------------------
// which = 0: Wc[tap][ci][coP]   (sp_corr:  destination channel = co)
// which = 1: Wt[tap][co][ciP]   (sp_corrT: destination channel = ci)
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int Co, int Ci, int k3,
                                    int which, int dP) {
    const int64_t total = (int64_t)k3 * (which == 0 ? Ci : Co) * dP;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int dch = (int)(i % dP);
        int64_t t = i / dP;
        float v = 0.f;
        if (which == 0) {
            const int ci = (int)(t % Ci), tap = (int)(t / Ci);
            if (dch < Co) v = w[((int64_t)dch * Ci + ci) * k3 + tap];
        } else {
            const int co = (int)(t % Co), tap = (int)(t / Co);
            if (dch < Ci) v = w[((int64_t)co * Ci + dch) * k3 + tap];
        }
        wp[i] = v;
    }
}

// ---- generic correlation: one thread = one O-side voxel x COT output channels ------------------------------------
template <int COT>
__global__ void __launch_bounds__(256)
corr_generic_kernel(SpConvDesc d, int nPerG, int coP, const float* __restrict__ src, const float* __restrict__ wp,
                    const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                    float* __restrict__ dst) {
    const int ncg = coP / COT;
    const int64_t total = (int64_t)d.N * d.Do * d.Ho * d.Wo * ncg;
    const bool vec = (d.Ci % 4 == 0) && (d.ldi % 4 == 0);
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % ncg);
        const int64_t v = idx / ncg;
        int64_t t = v;
        const int ow = (int)(t % d.Wo); t /= d.Wo;
        const int oh = (int)(t % d.Ho); t /= d.Ho;
        const int od = (int)(t % d.Do);
        const int n = (int)(t / d.Do);
        const int g = n / nPerG;
        const float* sc = scale ? scale + (int64_t)g * d.Ci : nullptr;
        const float* sh = scale ? shift + (int64_t)g * d.Ci : nullptr;
        float acc[COT];
#pragma unroll
        for (int j = 0; j < COT; ++j) acc[j] = 0.f;
        for (int kd = 0; kd < d.k; ++kd) {
            const int id = od * d.s - d.pd + kd;
            if (id < 0 || id >= d.Di) continue;
            for (int kh = 0; kh < d.k; ++kh) {
                const int ih = oh * d.s - d.ph + kh;
                if (ih < 0 || ih >= d.Hi) continue;
                for (int kw = 0; kw < d.k; ++kw) {
                    const int iw = ow * d.s - d.pw + kw;
                    if (iw < 0 || iw >= d.Wi) continue;
                    const int tap = (kd * d.k + kh) * d.k + kw;
                    const float* xp = src + ((((int64_t)n * d.Di + id) * d.Hi + ih) * d.Wi + iw) * d.ldi;
                    const float* wq = wp + (int64_t)tap * d.Ci * coP + cg * COT;
                    if (vec) {
                        for (int ci = 0; ci < d.Ci; ci += 4) {
                            float4 xv = *reinterpret_cast<const float4*>(xp + ci);
                            if (sc) {
                                xv.x = fmaf(xv.x, sc[ci + 0], sh[ci + 0]);
                                xv.y = fmaf(xv.y, sc[ci + 1], sh[ci + 1]);
                                xv.z = fmaf(xv.z, sc[ci + 2], sh[ci + 2]);
                                xv.w = fmaf(xv.w, sc[ci + 3], sh[ci + 3]);
                            }
                            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float* wr = wq + (int64_t)(ci + q) * coP;
#pragma unroll
                                for (int j = 0; j < COT; j += 4) {
                                    const float4 wv = *reinterpret_cast<const float4*>(wr + j);
                                    acc[j + 0] = fmaf(xs[q], wv.x, acc[j + 0]);
                                    acc[j + 1] = fmaf(xs[q], wv.y, acc[j + 1]);
                                    acc[j + 2] = fmaf(xs[q], wv.z, acc[j + 2]);
                                    acc[j + 3] = fmaf(xs[q], wv.w, acc[j + 3]);
                                }
                            }
                        }
                    } else {
                        for (int ci = 0; ci < d.Ci; ++ci) {
                            float xv = xp[ci];
                            if (sc) xv = fmaf(xv, sc[ci], sh[ci]);
                            const float* wr = wq + (int64_t)ci * coP;
#pragma unroll
                            for (int j = 0; j < COT; j += 4) {
                                const float4 wv = *reinterpret_cast<const float4*>(wr + j);
                                acc[j + 0] = fmaf(xv, wv.x, acc[j + 0]);
                                acc[j + 1] = fmaf(xv, wv.y, acc[j + 1]);
                                acc[j + 2] = fmaf(xv, wv.z, acc[j + 2]);
                                acc[j + 3] = fmaf(xv, wv.w, acc[j + 3]);
                            }
                        }
                    }
                }
            }
        }
        float* yp = dst + v * d.ldo;
#pragma unroll
        for (int j = 0; j < COT; ++j) {
            const int co = cg * COT + j;
            if (co < d.Co) {
                float r = acc[j] + (bias ? bias[co] : 0.f);
                yp[co] = sp_act_fwd(r, d.act, d.alpha);
            }
        }
    }
}

// ---- generic transposed correlation (gather form): one thread = one I-side voxel x CIT channels -------------------
template <int CIT>
__global__ void __launch_bounds__(256)
corrT_generic_kernel(SpConvDesc d, int nPerG, int ciP, const float* __restrict__ src, const float* __restrict__ wp,
                     const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                     float* __restrict__ dst) {
    const int ncg = ciP / CIT;
    const int64_t total = (int64_t)d.N * d.Di * d.Hi * d.Wi * ncg;
    const bool vec = (d.Co % 4 == 0) && (d.ldo % 4 == 0);
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int cg = (int)(idx % ncg);
        const int64_t v = idx / ncg;
        int64_t t = v;
        const int iw = (int)(t % d.Wi); t /= d.Wi;
        const int ih = (int)(t % d.Hi); t /= d.Hi;
        const int id = (int)(t % d.Di);
        const int n = (int)(t / d.Di);
        const int g = n / nPerG;
        const float* sc = scale ? scale + (int64_t)g * d.Co : nullptr;
        const float* sh = scale ? shift + (int64_t)g * d.Co : nullptr;
        float acc[CIT];
#pragma unroll
        for (int j = 0; j < CIT; ++j) acc[j] = 0.f;
        for (int kd = 0; kd < d.k; ++kd) {
            const int td = id + d.pd - kd;
            if (td < 0 || (td % d.s) != 0) continue;
            const int od = td / d.s;
            if (od >= d.Do) continue;
            for (int kh = 0; kh < d.k; ++kh) {
                const int th = ih + d.ph - kh;
                if (th < 0 || (th % d.s) != 0) continue;
                const int oh = th / d.s;
                if (oh >= d.Ho) continue;
                for (int kw = 0; kw < d.k; ++kw) {
                    const int tw = iw + d.pw - kw;
                    if (tw < 0 || (tw % d.s) != 0) continue;
                    const int ow = tw / d.s;
                    if (ow >= d.Wo) continue;
                    const int tap = (kd * d.k + kh) * d.k + kw;
                    const float* op = src + ((((int64_t)n * d.Do + od) * d.Ho + oh) * d.Wo + ow) * d.ldo;
                    const float* wq = wp + (int64_t)tap * d.Co * ciP + cg * CIT;
                    if (vec) {
                        for (int co = 0; co < d.Co; co += 4) {
                            float4 ov = *reinterpret_cast<const float4*>(op + co);
                            if (sc) {
                                ov.x = fmaf(ov.x, sc[co + 0], sh[co + 0]);
                                ov.y = fmaf(ov.y, sc[co + 1], sh[co + 1]);
                                ov.z = fmaf(ov.z, sc[co + 2], sh[co + 2]);
                                ov.w = fmaf(ov.w, sc[co + 3], sh[co + 3]);
                            }
                            const float os[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float* wr = wq + (int64_t)(co + q) * ciP;
#pragma unroll
                                for (int j = 0; j < CIT; j += 4) {
                                    const float4 wv = *reinterpret_cast<const float4*>(wr + j);
                                    acc[j + 0] = fmaf(os[q], wv.x, acc[j + 0]);
                                    acc[j + 1] = fmaf(os[q], wv.y, acc[j + 1]);
                                    acc[j + 2] = fmaf(os[q], wv.z, acc[j + 2]);
                                    acc[j + 3] = fmaf(os[q], wv.w, acc[j + 3]);
                                }
                            }
                        }
                    } else {
                        for (int co = 0; co < d.Co; ++co) {
                            float ov = op[co];
                            if (sc) ov = fmaf(ov, sc[co], sh[co]);
                            const float* wr = wq + (int64_t)co * ciP;
#pragma unroll
                            for (int j = 0; j < CIT; j += 4) {
                                const float4 wv = *reinterpret_cast<const float4*>(wr + j);
                                acc[j + 0] = fmaf(ov, wv.x, acc[j + 0]);
                                acc[j + 1] = fmaf(ov, wv.y, acc[j + 1]);
                                acc[j + 2] = fmaf(ov, wv.z, acc[j + 2]);
                                acc[j + 3] = fmaf(ov, wv.w, acc[j + 3]);
                            }
                        }
                    }
                }
            }
        }
        float* yp = dst + v * d.ldi;
#pragma unroll
        for (int j = 0; j < CIT; ++j) {
            const int ci = cg * CIT + j;
            if (ci < d.Ci) {
                float r = acc[j] + (bias ? bias[ci] : 0.f);
                yp[ci] = sp_act_fwd(r, d.act, d.alpha);
            }
        }
    }
}

// ---- generic weight gradient --------------------------------------------------------------------------------
// Work item = (tap, ci-quad, co-quad): a 4x4 register tile of dW.  grid.x covers the items, grid.y splits the
// O-side voxels into chunks; each block writes its partial dW into ws[chunk] and a second kernel reduces the
// chunks in fixed order (deterministic, no float atomics).
struct WgradPlan {
    int ciQ, coQ, items, ipb, vl, blocks_x, chunks;
    int64_t ov, per_chunk, wn;
};

WgradPlan wgrad_plan(const SpConvDesc* d) {
    WgradPlan p;
    p.ciQ = (d->Ci + 3) / 4;
    p.coQ = (d->Co + 3) / 4;
    const int k3 = d->k * d->k * d->k;
    p.items = k3 * p.ciQ * p.coQ;
    // few work items (1x1x1 layers: 16 or 4): split the voxels of a chunk over `vl` lanes per item so that the CTA
    // still has 256 busy threads; the lanes are summed through shared memory at the end
    p.ipb = p.items < 256 ? p.items : 256;
    p.vl = 1;
    while (p.vl * 2 * p.ipb <= 256) p.vl *= 2;
    p.blocks_x = (p.items + p.ipb - 1) / p.ipb;
    p.ov = (int64_t)d->N * d->Do * d->Ho * d->Wo;
    p.wn = (int64_t)d->Co * d->Ci * k3;
    int64_t chunks = sp_cdiv(4 * 148, p.blocks_x);                 // ~4 CTAs per SM in flight
    const int64_t min_vox = 32 * p.vl;
    chunks = chunks < sp_cdiv(p.ov, min_vox) ? chunks : sp_cdiv(p.ov, min_vox);
    const int64_t cap = (int64_t)(96u << 20) / (p.wn * 4);         // keep partials under 96 MB
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    p.chunks = (int)chunks;
    p.per_chunk = sp_cdiv(p.ov, p.chunks);
    return p;
}

template <bool VEC_I, bool VEC_O>
__global__ void __launch_bounds__(256)
wgrad_generic_kernel(SpConvDesc d, int nPerG, int ciQ, int coQ, int items, int ipb, int vl, int64_t ov, int64_t per_chunk,
                     const float* __restrict__ iside, const float* __restrict__ i_scale, const float* __restrict__ i_shift,
                     const float* __restrict__ oside, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                     float* __restrict__ ws) {
    __shared__ float red[256 * 16];
    const int il = threadIdx.x % ipb;           // item within the CTA
    const int lane_v = threadIdx.x / ipb;       // voxel lane (0 .. vl-1; threads beyond ipb*vl idle)
    const int item = blockIdx.x * ipb + il;
    const bool live = (item < items) && (lane_v < vl);
    const int it = live ? item : 0;
    const int coq = it % coQ;
    const int ciq = (it / coQ) % ciQ;
    const int tap = it / (coQ * ciQ);
    const int kw = tap % d.k, kh = (tap / d.k) % d.k, kd = tap / (d.k * d.k);
    const int ci0 = ciq * 4, co0 = coq * 4;
    const int k3 = d.k * d.k * d.k;

    float acc[4][4];  // [co][ci]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    const int64_t v0 = (int64_t)blockIdx.y * per_chunk;
    const int64_t v1 = (v0 + per_chunk < ov) ? v0 + per_chunk : ov;
    if (live) {
        for (int64_t v = v0 + lane_v; v < v1; v += vl) {
            int64_t t = v;
            const int ow = (int)(t % d.Wo); t /= d.Wo;
            const int oh = (int)(t % d.Ho); t /= d.Ho;
            const int od = (int)(t % d.Do);
            const int n = (int)(t / d.Do);
            const int id = od * d.s - d.pd + kd, ih = oh * d.s - d.ph + kh, iw = ow * d.s - d.pw + kw;
            if (id < 0 || id >= d.Di || ih < 0 || ih >= d.Hi || iw < 0 || iw >= d.Wi) continue;
            const int g = n / nPerG;
            const float* ip = iside + ((((int64_t)n * d.Di + id) * d.Hi + ih) * d.Wi + iw) * d.ldi + ci0;
            const float* op = oside + v * d.ldo + co0;
            float xi[4], xo[4];
            if (VEC_I) {
                const float4 a = *reinterpret_cast<const float4*>(ip);
                xi[0] = a.x; xi[1] = a.y; xi[2] = a.z; xi[3] = a.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) xi[q] = (ci0 + q < d.Ci) ? ip[q] : 0.f;
            }
            if (VEC_O) {
                const float4 a = *reinterpret_cast<const float4*>(op);
                xo[0] = a.x; xo[1] = a.y; xo[2] = a.z; xo[3] = a.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) xo[q] = (co0 + q < d.Co) ? op[q] : 0.f;
            }
            if (i_scale) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (ci0 + q < d.Ci) xi[q] = fmaf(xi[q], i_scale[g * d.Ci + ci0 + q], i_shift[g * d.Ci + ci0 + q]);
            }
            if (o_scale) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (co0 + q < d.Co) xo[q] = fmaf(xo[q], o_scale[g * d.Co + co0 + q], o_shift[g * d.Co + co0 + q]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(xo[a], xi[b], acc[a][b]);
        }
    }
    // sum the voxel lanes of each item (fixed order) and write the CTA's partial
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) red[threadIdx.x * 16 + a * 4 + b] = acc[a][b];
    __syncthreads();
    if (lane_v == 0 && item < items) {
        float* wsp = ws + (int64_t)blockIdx.y * ((int64_t)d.Co * d.Ci * k3);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                float s = acc[a][b];
                for (int l = 1; l < vl; ++l) s += red[(l * ipb + il) * 16 + a * 4 + b];
                if (co0 + a < d.Co && ci0 + b < d.Ci) wsp[((int64_t)(co0 + a) * d.Ci + (ci0 + b)) * k3 + tap] = s;
            }
    }
}

// ---- bias gradient: column sums of a [rows][ld] matrix -----------------------------------------------------------
__global__ void __launch_bounds__(256)
bias_grad_partial_kernel(const float* __restrict__ g, int64_t rows, int C, int ld, int64_t rows_per_block, double* __restrict__ acc) {
    // thread (r, c): c = tid % C-lanes; accumulate in fp64, block reduce through shared memory, fp64 atomics
    extern __shared__ double sm[];
    const int lanesC = C < 256 ? C : 256;
    const int R = 256 / lanesC;
    const int c_l = threadIdx.x % lanesC, r_l = threadIdx.x / lanesC;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
    for (int cb = 0; cb < C; cb += lanesC) {   // uniform trip count: every thread reaches the barriers below
        const int c = cb + c_l;
        double s = 0.0;
        if (r_l < R && c < C)
            for (int64_t r = r0 + r_l; r < r1; r += R) s += (double)g[r * ld + c];
        sm[threadIdx.x] = s;
        __syncthreads();
        if (r_l == 0 && c < C) {
            for (int q = 1; q < R; ++q) s += sm[q * lanesC + c_l];
            atomicAdd(&acc[c], s);
        }
        __syncthreads();
    }
}
__global__ void bias_grad_final_kernel(const double* __restrict__ acc, int C, float* __restrict__ db, float beta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) db[c] = (beta == 0.f) ? (float)acc[c] : fmaf(beta, db[c], (float)acc[c]);
}

int grid_for(int64_t work_items, int threads = 256, int waves = 16) {
    int64_t b = sp_cdiv(work_items, threads);
    const int64_t cap = (int64_t)sp_num_sms() * waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

// ---- tensor-core tier plumbing ------------------------------------------------------------------------------------------
// The packed weight buffer of a layer the tcgen05 tier serves is [fp32 FFMA layout | bf16-split UMMA image]: the image
// travels with the fp32 pack so the C-ABI (and its host-side per-parameter-version cache) stays unchanged.
namespace {

size_t ffma_packed_floats(const SpConvDesc* d, int which) {
    const int k3 = d->k * d->k * d->k;
    return which == 0 ? (size_t)k3 * d->Ci * round_up(d->Co, kPad) : (size_t)k3 * d->Co * round_up(d->Ci, kPad);
}

// geometry seen by the correlation kernels when sp_corrT (stride 1) is run as a flipped correlation
SpConvDesc flipped_desc(const SpConvDesc* d) {
    SpConvDesc f = *d;
    f.Di = d->Do; f.Hi = d->Ho; f.Wi = d->Wo; f.Ci = d->Co; f.ldi = d->ldo;
    f.Do = d->Di; f.Ho = d->Hi; f.Wo = d->Wi; f.Co = d->Ci; f.ldo = d->ldi;
    f.pd = d->k - 1 - d->pd; f.ph = d->k - 1 - d->ph; f.pw = d->k - 1 - d->pw;
    return f;
}

bool tc_serves(const SpConvDesc* d, int which, SpTcCfg* cfg) {
    if (which == 0) return sp_tc_corr_supported(d, cfg);
    if (d->s != 1) return false;
    const SpConvDesc f = flipped_desc(d);
    return f.pd >= 0 && f.ph >= 0 && f.pw >= 0 && sp_tc_corr_supported(&f, cfg);
}

// which op of a layer the GEMM tier (sp_conv_gemm.cuh) serves: everything wide enough that no spatial tier takes
bool gemm_corr(const SpConvDesc* d) { return !sp_pw_fwd_supported(d, d->Ci) && sp_gemm_serves(d); }
bool gemm_corrT(const SpConvDesc* d) {
    if (sp_gemm_disabled() || d->k < 2) return false;
    if (d->s == 1) {
        const SpConvDesc f = flipped_desc(d);
        if (sp_thin_bwd_supported(&f)) return false;
        if (f.pd >= 0 && f.ph >= 0 && f.pw >= 0 && sp_tiled_corr_supported(&f)) return false;
    }
    if (sp_tiledT_supported(d) || sp_k2s2_supported(d)) return false;
    return d->Ci * d->k * d->k * d->k >= 256 && d->Co >= 16 && d->Ci >= 8;
}
bool gemm_wgrad(const SpConvDesc* d) {
    if (sp_gemm_disabled() || d->k < 2 || sp_tiled_wgrad_supported(d) || sp_k2s2_supported(d)) return false;
    return d->Ci * d->k * d->k * d->k >= 256 && d->Co >= 16 && d->Ci >= 8;
}

// ONE tier choice per (geometry, direction), shared by sp_packed_weight_floats / sp_pack_weights (which layout to write) and
// sp_corr / sp_corrT (which kernel reads it), so the two can never disagree — whatever tiers are switched off.
enum SpTier { TIER_PW, TIER_TC, TIER_K2S2, TIER_THIN, TIER_TILED, TIER_TILEDT, TIER_GEMM, TIER_GENERIC };

SpTier corr_tier(const SpConvDesc* d, SpTcCfg* cfg) {
    if (sp_pw_fwd_supported(d, d->Ci)) return TIER_PW;
    if (tc_serves(d, 0, cfg)) return TIER_TC;
    if (sp_k2s2_supported(d)) return TIER_K2S2;          // (falls to the tiers below at run time if a pointer is misaligned)
    if (sp_thin_fwd_supported(d)) return TIER_THIN;
    if (sp_tiled_corr_supported(d)) return TIER_TILED;
    if (gemm_corr(d)) return TIER_GEMM;
    return TIER_GENERIC;
}

SpTier corrT_tier(const SpConvDesc* d, SpTcCfg* cfg) {
    if (sp_pw_fwd_supported(d, d->Co)) return TIER_PW;
    if (d->s == 1) {
        const SpConvDesc f = flipped_desc(d);
        if (tc_serves(d, 1, cfg)) return TIER_TC;
        if (sp_thin_bwd_supported(&f)) return TIER_THIN;
        if (f.pd >= 0 && f.ph >= 0 && f.pw >= 0 && sp_tiled_corr_supported(&f)) return TIER_TILED;
    }
    if (sp_k2s2_supported(d)) return TIER_K2S2;
    if (sp_tiledT_supported(d)) return TIER_TILEDT;
    if (gemm_corrT(d)) return TIER_GEMM;
    return TIER_GENERIC;
}

// Which kernel generation serves a layer of the tensor-core tier: generation 3 (kw-stacked N, rolling depth window, TMA) for output
// widths up to 16 and for every slice / pass of the wide layers; for the 17..24-wide layers of the CAE's second level (kw-stacked N =
// 240 columns, 2 accumulators, a half-empty second K pass) it depends on the direction.  The bf16 mode exists in generation 3 only.
// which = 0: sp_corr (forward of a Conv3d), 1: sp_corrT (its dgrad).  Measured in the CAE step after the elected-issue / TMA work:
// the 24-wide FORWARDS are faster on generation 3 (0.795 -> 0.654, 0.527 -> 0.488, 0.518 -> 0.482 ms), their dgrads on generation 2
// (0.488 vs 0.588, 0.381 vs 0.445 ms).
bool use_gen3(const SpTcCfg& cfg, int which) {
    static int g2 = -1;   // SP_TC3_COP24=0: generation 2 for every 17..24-wide layer (A/B checks)
    if (g2 < 0) {
        const char* e = getenv("SP_TC3_COP24");
        g2 = (e && e[0] == '0') ? 1 : 0;
    }
    return sp_tc_terms() == 1 || (sp_tc_terms() == 5 && (cfg.cop == 16 || (which == 0 && !g2)));
}

int tc_corr_launch(const SpConvDesc* d, const SpTcCfg& cfg, int which, int nPerG, const float* src, const float* wimg, const float* bias,
                   const float* scale, const float* shift, float* dst, cudaStream_t st) {
    const uint4* img = reinterpret_cast<const uint4*>(wimg);
    if (use_gen3(cfg, which)) return sp_tc3_corr_launch(d, nPerG, sp_tc_image_terms(), src, img, bias, scale, shift, dst, st);
    if (sp_tc_terms() == 4 || sp_tc_terms() == 5) return sp_tc2_corr_launch(d, nPerG, src, img, bias, scale, shift, dst, st);
    if (sp_tc_terms() == 2) return sp_tc_corr_launch_t<16, 16, 2, 4>(d, nPerG, src, img, bias, scale, shift, dst, st);
    return sp_tc_corr_launch_t<16, 16, 3, 2>(d, nPerG, src, img, bias, scale, shift, dst, st);
}

}  // namespace

// =============================================================================================================
extern "C" {

int sp_get_tc_terms(void) { return sp_tc_terms(); }

int sp_set_tc_terms(int terms) {
    SP_REQUIRE(terms >= 0 && terms <= 5, "sp_set_tc_terms: mode must be 0 (off) .. 5, got %d", terms);
    sp_tc_terms_ref() = terms;
    return 0;
}

int sp_set_wgrad_tc_options(int generation, int max_ctas) {
    SP_REQUIRE(generation == 1 || generation == 2, "sp_set_wgrad_tc_options: generation must be 1 or 2, got %d", generation);
    SP_REQUIRE(max_ctas >= 0, "sp_set_wgrad_tc_options: max_ctas must be >= 0, got %d", max_ctas);
    sp_wtc4_generation_ref() = generation;
    sp_wtc4_grid_cap_ref() = max_ctas;
    return 0;
}

size_t sp_packed_weight_floats(const SpConvDesc* d, int which) {
    if (!d) return 0;
    size_t n = ffma_packed_floats(d, which);
    SpTcCfg cfg;
    const bool tc = (which == 0 ? corr_tier(d, &cfg) : corrT_tier(d, &cfg)) == TIER_TC;
    if (tc && use_gen3(cfg, which))
        n += (size_t)cfg.passes * cfg.nslices * sp_tc3_wimg_u4(cfg.cop, sp_tc_image_terms()) * 4;
    else if (tc) n += (size_t)cfg.passes * cfg.nslices * sp_tc_wimg_bytes(cfg.cip, cfg.cop, sp_tc_image_terms()) / sizeof(float);
    return n;
}

int sp_pack_weights(const SpConvDesc* d, int which, const float* w_torch, float* w_packed, void* stream) {
    if (int e = check_desc(d)) return e;
    SP_REQUIRE(which == 0 || which == 1, "sp_pack_weights: which must be 0 (corr) or 1 (corrT)");
    SP_REQUIRE(w_torch && w_packed, "sp_pack_weights: NULL pointer");
    const int k3 = d->k * d->k * d->k;
    const int dP = round_up(which == 0 ? d->Co : d->Ci, kPad);
    const int64_t total = (int64_t)ffma_packed_floats(d, which);
    SpTcCfg cfg;
    const SpTier tier = which == 0 ? corr_tier(d, &cfg) : corrT_tier(d, &cfg);
    if (which == 1 && tier == TIER_GEMM) {   // GEMM tier: Wt[co][tap][ciP] (same size as the [tap][co][ciP] layout)
        sp_gemm::pack_wt_gemm_kernel<<<grid_for(total), 256, 0, sp_stream(stream)>>>(w_torch, w_packed, d->Co, d->Ci, k3, dP);
        SP_LAUNCH_OK("pack_wt_gemm_kernel");
        return 0;
    }
    const bool tc = tier == TIER_TC;
    if (!tc) {      // a layer the tensor-core tier serves never reads the FFMA layout (same predicate at pack and at run time)
        pack_weights_kernel<<<grid_for(total), 256, 0, sp_stream(stream)>>>(w_torch, w_packed, d->Co, d->Ci, k3, which, dP);
        SP_LAUNCH_OK("pack_weights_kernel");
    }
    if (tc && use_gen3(cfg, which))
        return sp_tc3_pack_launch(d, which, sp_tc_image_terms(), cfg.cop, w_torch, w_packed + total, sp_stream(stream), cfg.passes, cfg.nslices);
    if (tc)
        return sp_tc_pack_launch(d, which, sp_tc_image_terms(), cfg.cip, cfg.cop, w_torch, w_packed + total, sp_stream(stream), cfg.passes, cfg.nslices);
    return 0;
}

size_t sp_conv_workspace_bytes(const SpConvDesc* d, int which) {
    if (!d || d->N <= 0) return 0;
    SpTcCfg cfg;
    if (which == 0) return corr_tier(d, &cfg) == TIER_GEMM ? sp_gemm_corr_ws_bytes(d) + 256 : 0;
    return corrT_tier(d, &cfg) == TIER_GEMM ? sp_gemm_corrT_ws_bytes(d) + 256 : 0;
}

int sp_corr(const SpConvDesc* d, const float* src, const float* wp, const float* bias, const float* scale,
            const float* shift, int G, float* dst, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_desc(d)) return e;
    SP_REQUIRE(src && wp && dst, "sp_corr: NULL pointer");
    SP_REQUIRE((scale == nullptr) == (shift == nullptr), "sp_corr: scale and shift must be given together");
    SP_REQUIRE(G >= 1 && d->N % G == 0, "sp_corr: N=%d not divisible by G=%d", d->N, G);
    const int coP = round_up(d->Co, kPad);
    const int nPerG = d->N / G;
    if (sp_pw_fwd_supported(d, d->Ci)) {
        const int64_t vox = (int64_t)d->Do * d->Ho * d->Wo;
        return sp_pw_fwd_launch(src, d->ldi, d->Ci, dst, d->ldo, d->Co, (int64_t)d->N * vox, (int64_t)nPerG * vox, wp, bias, scale, shift,
                                d->act, d->alpha, sp_stream(stream));
    }
    SpTcCfg cfg;
    const SpTier tier = corr_tier(d, &cfg);
    if (tier == TIER_TC)
        return tc_corr_launch(d, cfg, 0, nPerG, src, wp + ffma_packed_floats(d, 0), bias, scale, shift, dst, sp_stream(stream));
    if (sp_k2s2_supported(d) && sp_k2s2_aligned(src, dst) && (!bias || sp_k2s2_aligned(bias, wp)))
        return sp_k2s2_down_launch(d, nPerG, src, wp, bias, scale, shift, dst, sp_stream(stream));
    if (sp_thin_fwd_supported(d)) return sp_thin_fwd_launch(d, nPerG, src, wp, /*flip=*/0, bias, scale, shift, dst, sp_stream(stream));
    if (sp_tiled_corr_supported(d)) return sp_tiled_corr_launch(d, nPerG, src, wp, /*flip=*/0, bias, scale, shift, dst, sp_stream(stream));
    if (tier == TIER_GEMM) {
        SP_REQUIRE(ws && ws_bytes >= sp_conv_workspace_bytes(d, 0), "sp_corr: workspace too small (%zu < %zu)", ws_bytes,
                   sp_conv_workspace_bytes(d, 0));
        return sp_gemm_corr_launch(d, nPerG, src, wp, bias, scale, shift, dst, (float*)ws, sp_stream(stream));
    }
    const int64_t work = (int64_t)d->N * d->Do * d->Ho * d->Wo * (coP / 8);
    corr_generic_kernel<8><<<grid_for(work), 256, 0, sp_stream(stream)>>>(*d, nPerG, coP, src, wp, bias, scale, shift, dst);
    SP_LAUNCH_OK("corr_generic_kernel");
    return 0;
}

int sp_corrT(const SpConvDesc* d, const float* src, const float* wp, const float* bias, const float* scale,
             const float* shift, int G, float* dst, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_desc(d)) return e;
    SP_REQUIRE(src && wp && dst, "sp_corrT: NULL pointer");
    SP_REQUIRE((scale == nullptr) == (shift == nullptr), "sp_corrT: scale and shift must be given together");
    SP_REQUIRE(G >= 1 && d->N % G == 0, "sp_corrT: N=%d not divisible by G=%d", d->N, G);
    const int ciP = round_up(d->Ci, kPad);
    const int nPerG = d->N / G;
    if (sp_pw_fwd_supported(d, d->Co)) {   // 1x1x1: the transposed correlation is the same channel mix on Wt[0][co][ciP]
        const int64_t vox = (int64_t)d->Do * d->Ho * d->Wo;
        return sp_pw_fwd_launch(src, d->ldo, d->Co, dst, d->ldi, d->Ci, (int64_t)d->N * vox, (int64_t)nPerG * vox, wp, bias, scale, shift,
                                d->act, d->alpha, sp_stream(stream));
    }
    SpTcCfg cfg;
    const SpTier tier = corrT_tier(d, &cfg);
    if (d->s == 1) {
        // stride 1: the transposed correlation is a correlation with flipped taps, swapped channel roles and
        // padding k-1-p; Wt[tap][co][ciP] read with flipped tap index is exactly that correlation's Wc.
        const SpConvDesc f = flipped_desc(d);
        if (tier == TIER_TC)
            return tc_corr_launch(&f, cfg, 1, nPerG, src, wp + ffma_packed_floats(d, 1), bias, scale, shift, dst, sp_stream(stream));
        if (sp_thin_bwd_supported(&f)) return sp_thin_bwd_launch(&f, nPerG, src, wp, bias, scale, shift, dst, sp_stream(stream));
        if (f.pd >= 0 && f.ph >= 0 && f.pw >= 0 && sp_tiled_corr_supported(&f))
            return sp_tiled_corr_launch(&f, nPerG, src, wp, /*flip=*/1, bias, scale, shift, dst, sp_stream(stream));
    }
    if (sp_k2s2_supported(d) && sp_k2s2_aligned(src, dst) && (!bias || sp_k2s2_aligned(bias, wp)))
        return sp_k2s2_up_launch(d, nPerG, src, wp, bias, scale, shift, dst, sp_stream(stream));
    if (sp_tiledT_supported(d)) return sp_tiledT_launch(d, nPerG, src, wp, bias, scale, shift, dst, sp_stream(stream));
    if (tier == TIER_GEMM) {
        SP_REQUIRE(ws && ws_bytes >= sp_conv_workspace_bytes(d, 1), "sp_corrT: workspace too small (%zu < %zu)", ws_bytes,
                   sp_conv_workspace_bytes(d, 1));
        return sp_gemm_corrT_launch(d, nPerG, src, wp, bias, scale, shift, dst, (float*)ws, sp_stream(stream));
    }
    const int64_t work = (int64_t)d->N * d->Di * d->Hi * d->Wi * (ciP / 8);
    corrT_generic_kernel<8><<<grid_for(work), 256, 0, sp_stream(stream)>>>(*d, nPerG, ciP, src, wp, bias, scale, shift, dst);
    SP_LAUNCH_OK("corrT_generic_kernel");
    return 0;
}

size_t sp_wgrad_workspace_bytes(const SpConvDesc* d) {
    if (!d || d->N <= 0) return 0;
    size_t generic = 0, tiled = 0;
    {
        const WgradPlan p = wgrad_plan(d);
        generic = (size_t)p.chunks * p.wn * sizeof(float);
    }
    tiled = sp_tiled_wgrad_workspace_bytes(d);
    const size_t pw = sp_pw_wgrad_workspace_bytes(d);
    if (pw > tiled) tiled = pw;
    if (sp_tc_wgrad_workspace_bytes(d) > tiled) tiled = sp_tc_wgrad_workspace_bytes(d);
    if (sp_tc24_wgrad_workspace_bytes(d) > tiled) tiled = sp_tc24_wgrad_workspace_bytes(d);
    if (sp_tc_wgrad_sliced_workspace_bytes(d) > tiled) tiled = sp_tc_wgrad_sliced_workspace_bytes(d);
    if ((d->k == 3 || d->k == 2) && d->s == 2 && d->Ci > 8 && d->Ci <= 32 && sp_tc4s2_wgrad_workspace_bytes(d) > tiled) tiled = sp_tc4s2_wgrad_workspace_bytes(d);
    if (d->k == 3 && d->s == 1 && d->Ci >= 2 && d->Co > 8 && d->Ci <= 96 && d->Co <= 64 && sp_tc4_wgrad_workspace_bytes(d) > tiled)
        tiled = sp_tc4_wgrad_workspace_bytes(d);      // second-generation tcgen05 weight gradient (any G)
    if (sp_thin_wgrad_workspace_bytes(d) > tiled) tiled = sp_thin_wgrad_workspace_bytes(d);
    if (sp_k2s2_wgrad_workspace_bytes(d) > tiled) tiled = sp_k2s2_wgrad_workspace_bytes(d);
    if (gemm_wgrad(d) && sp_gemm_wgrad_ws_bytes(d) > tiled) tiled = sp_gemm_wgrad_ws_bytes(d);
    return (generic > tiled ? generic : tiled) + 256;
}

int sp_wgrad(const SpConvDesc* d, const float* iside, const float* i_scale, const float* i_shift, const float* oside,
             const float* o_scale, const float* o_shift, int G, float* dw, float beta, void* ws, size_t ws_bytes,
             void* stream) {
    if (int e = check_desc(d)) return e;
    SP_REQUIRE(iside && oside && dw && ws, "sp_wgrad: NULL pointer");
    SP_REQUIRE((i_scale == nullptr) == (i_shift == nullptr) && (o_scale == nullptr) == (o_shift == nullptr),
               "sp_wgrad: scale and shift must be given together");
    SP_REQUIRE(G >= 1 && d->N % G == 0, "sp_wgrad: N=%d not divisible by G=%d", d->N, G);
    SP_REQUIRE(ws_bytes >= sp_wgrad_workspace_bytes(d), "sp_wgrad: workspace too small (%zu < %zu)", ws_bytes,
               sp_wgrad_workspace_bytes(d));
    const int nPerG = d->N / G;
    if (sp_pw_wgrad_supported(d))
        return sp_pw_wgrad_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    if (sp_tc4s2_wgrad_supported(d, G) && ((reinterpret_cast<uintptr_t>(iside) | reinterpret_cast<uintptr_t>(oside)) & 15) == 0)
        return sp_tc4s2_wgrad_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    if (sp_k2s2_supported(d) && sp_k2s2_aligned(iside, oside))
        return sp_k2s2_wgrad_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    if (d->Ci >= 2 && sp_tc4_wgrad_supported(d, G) && sp_tc4_wgrad_aligned(d, iside, oside))
        return sp_tc4_wgrad_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    if (sp_thin_wgrad_supported(d))
        return sp_thin_wgrad_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    const bool al16 = ((reinterpret_cast<uintptr_t>(iside) | reinterpret_cast<uintptr_t>(oside)) & 15) == 0;
    if (sp_tc_wgrad_supported(d))
        return sp_tc_wgrad_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    if (sp_tc_wgrad_sliced_supported(d) && ((reinterpret_cast<uintptr_t>(iside) | reinterpret_cast<uintptr_t>(oside)) & 15) == 0)
        return sp_tc_wgrad_sliced_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    if (sp_tc24_wgrad_supported(d))
        return sp_tc24_wgrad_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    if (sp_tiled_wgrad_supported(d))
        return sp_tiled_wgrad_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    if (gemm_wgrad(d) && !sp_pw_wgrad_supported(d))
        return sp_gemm_wgrad_launch(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, (float*)ws, sp_stream(stream));
    const WgradPlan p = wgrad_plan(d);
    const bool vi = (d->Ci % 4 == 0) && (d->ldi % 4 == 0);
    const bool vo = (d->Co % 4 == 0) && (d->ldo % 4 == 0);
    dim3 grid(p.blocks_x, p.chunks);
    float* wsf = (float*)ws;
#define SP_WG(VI, VO)                                                                                             \
    wgrad_generic_kernel<VI, VO><<<grid, 256, 0, sp_stream(stream)>>>(*d, nPerG, p.ciQ, p.coQ, p.items, p.ipb, p.vl, \
                                                                      p.ov, p.per_chunk, iside, i_scale, i_shift,   \
                                                                      oside, o_scale, o_shift, wsf)
    if (vi && vo) SP_WG(true, true);
    else if (vi) SP_WG(true, false);
    else if (vo) SP_WG(false, true);
    else SP_WG(false, false);
#undef SP_WG
    SP_LAUNCH_OK("wgrad_generic_kernel");
    wgrad_reduce_kernel<<<grid_for(p.wn), 256, 0, sp_stream(stream)>>>(wsf, p.chunks, p.wn, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}

int sp_bias_grad(const float* g, int64_t rows, int C, int ld, float* db, float beta, double* acc, void* stream) {
    SP_REQUIRE(g && db && rows > 0 && C > 0 && ld >= C, "sp_bias_grad: bad arguments");
    SP_REQUIRE(C <= 4096, "sp_bias_grad: C=%d too large", C);
    SP_REQUIRE(acc != nullptr, "sp_bias_grad: workspace of C doubles required");
    SP_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * C, sp_stream(stream)));
    int64_t blocks = sp_cdiv(rows, 64);
    const int64_t cap = (int64_t)sp_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const int64_t rpb = sp_cdiv(rows, blocks);
    blocks = sp_cdiv(rows, rpb);
    bias_grad_partial_kernel<<<(int)blocks, 256, 256 * sizeof(double), sp_stream(stream)>>>(g, rows, C, ld, rpb, acc);
    SP_LAUNCH_OK("bias_grad_partial_kernel");
    bias_grad_final_kernel<<<(C + 255) / 256, 256, 0, sp_stream(stream)>>>(acc, C, db, beta);
    SP_LAUNCH_OK("bias_grad_final_kernel");
    return 0;
}

}  // extern "C"
