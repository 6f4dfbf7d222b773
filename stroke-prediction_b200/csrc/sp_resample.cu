// sp_resample.cu — U-Net resampling ops and layout glue (fp32, NDHWC).
//   MaxPool3d(2,2)                      common/model/Unet3D.py:39,41
//   Upsample(scale_factor=2, trilinear) common/model/Unet3D.py:44,46
//   centre-crop + channel concat        common/model/Unet3D.py:6-11,66-67,71-72
// All of these are HBM-bound: threads walk the channel axis fastest so every warp touches contiguous memory, and
// each element is read/written exactly once.
#include "sp_common.cuh"

namespace {

inline int ew_grid(int64_t n) {
    int64_t b = sp_cdiv(n, 256);
    const int64_t cap = (int64_t)sp_num_sms() * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---- max pool ------------------------------------------------------------------------------------------------
// floor mode: odd trailing planes are dropped
__global__ void __launch_bounds__(256)
maxpool2_fwd_kernel(const float* __restrict__ x, int N, int D, int H, int W, int C, int ldx, float* __restrict__ y, int ldy) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    const int64_t total = (int64_t)N * Do * Ho * Wo * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ow = (int)(t % Wo); t /= Wo;
        const int oh = (int)(t % Ho); t /= Ho;
        const int od = (int)(t % Do);
        const int n = (int)(t / Do);
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float v = x[((((int64_t)n * D + 2 * od + a) * H + 2 * oh + b) * W + 2 * ow + e) * ldx + c];
                    if (v > m || v != v) m = v;   // strict '>' keeps the first maximum; NaN propagates like ATen
                }
        y[(i / C) * ldy + c] = m;
    }
}

// gx is dense (ld = C) and fully written: the first element (d->h->w scan order) equal to the window maximum
// receives gy, everything else — including dropped odd planes — receives 0.
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy, int N, int D,
                    int H, int W, int C, float* __restrict__ gx) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2;
    const int64_t total = (int64_t)N * D * H * W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H); t /= H;
        const int d = (int)(t % D);
        const int n = (int)(t / D);
        const int od = d >> 1, oh = h >> 1, ow = w >> 1;
        float r = 0.f;
        if (od < Do && oh < Ho && ow < Wo) {
            const int64_t o = ((((int64_t)n * Do + od) * Ho + oh) * Wo + ow) * C + c;
            const float m = y[o];
            const float v = x[i];
            if (v == m) {
                // am I the first element of the window holding the maximum?
                bool first = true;
                const int me = ((d & 1) << 2) | ((h & 1) << 1) | (w & 1);
                for (int q = 0; q < me; ++q) {
                    const int a = q >> 2, b = (q >> 1) & 1, e = q & 1;
                    const float u = x[((((int64_t)n * D + 2 * od + a) * H + 2 * oh + b) * W + 2 * ow + e) * C + c];
                    if (u == m) { first = false; break; }
                }
                if (first) r = gy[o];
            }
        }
        gx[i] = r;
    }
}

// ---- trilinear x2 upsample -------------------------------------------------------------------------------------
// ATen source-index rule (UpSample.h area_pixel_compute_source_index):
//   align_corners = 0: src = max(0, (dst + 0.5) * 0.5 - 0.5)          (scale_factor given -> scale = 1/2)
//   align_corners = 1: src = dst * (in - 1) / (out - 1)
// i0 = floor(src), i1 = min(i0 + 1, in - 1), l1 = src - i0, l0 = 1 - l1.
__device__ __forceinline__ void up_src(int dst, int in, int out, int align, int& i0, int& i1, float& l0, float& l1) {
    float src;
    if (align) {
        const float sc = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
        src = sc * dst;
    } else {
        src = 0.5f * (dst + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
    }
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256)
upsample2_fwd_kernel(const float* __restrict__ x, int N, int D, int H, int W, int C, int ldx, float* __restrict__ y,
                     int ldy, int align) {
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
    const int64_t total = (int64_t)N * Do * Ho * Wo * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int ow = (int)(t % Wo); t /= Wo;
        const int oh = (int)(t % Ho); t /= Ho;
        const int od = (int)(t % Do);
        const int n = (int)(t / Do);
        int d0, d1, h0, h1, w0, w1;
        float ld0, ld1, lh0, lh1, lw0, lw1;
        up_src(od, D, Do, align, d0, d1, ld0, ld1);
        up_src(oh, H, Ho, align, h0, h1, lh0, lh1);
        up_src(ow, W, Wo, align, w0, w1, lw0, lw1);
        const float* xb = x + (int64_t)n * D * H * W * ldx + c;
#define XV(dd, hh, ww) xb[(((int64_t)(dd) * H + (hh)) * W + (ww)) * ldx]
        // same association order as ATen's upsample_trilinear3d CPU kernel
        const float r = ld0 * (lh0 * (lw0 * XV(d0, h0, w0) + lw1 * XV(d0, h0, w1)) + lh1 * (lw0 * XV(d0, h1, w0) + lw1 * XV(d0, h1, w1))) +
                        ld1 * (lh0 * (lw0 * XV(d1, h0, w0) + lw1 * XV(d1, h0, w1)) + lh1 * (lw0 * XV(d1, h1, w0) + lw1 * XV(d1, h1, w1)));
#undef XV
        y[(i / C) * ldy + c] = r;
    }
}

// Gather form of the transpose: input voxel i collects lambda-weighted gy from every output voxel whose stencil
// touches it (outputs 2i-3 .. 2i+3 cover both align_corners modes).  Deterministic, no atomics.
__global__ void __launch_bounds__(256)
upsample2_bwd_kernel(const float* __restrict__ gy, int ldgy, int N, int D, int H, int W, int C, float* __restrict__ gx,
                     int ldgx, int align) {
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
    const int64_t total = (int64_t)N * D * H * W * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H); t /= H;
        const int d = (int)(t % D);
        const int n = (int)(t / D);
        float wd[7], wh[7], ww[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            int i0, i1;
            float l0, l1;
            int o = 2 * d - 3 + q;
            wd[q] = 0.f;
            if (o >= 0 && o < Do) {
                up_src(o, D, Do, align, i0, i1, l0, l1);
                wd[q] = (i0 == d ? l0 : 0.f) + (i1 == d ? l1 : 0.f);
            }
            o = 2 * h - 3 + q;
            wh[q] = 0.f;
            if (o >= 0 && o < Ho) {
                up_src(o, H, Ho, align, i0, i1, l0, l1);
                wh[q] = (i0 == h ? l0 : 0.f) + (i1 == h ? l1 : 0.f);
            }
            o = 2 * w - 3 + q;
            ww[q] = 0.f;
            if (o >= 0 && o < Wo) {
                up_src(o, W, Wo, align, i0, i1, l0, l1);
                ww[q] = (i0 == w ? l0 : 0.f) + (i1 == w ? l1 : 0.f);
            }
        }
        const float* gb = gy + (int64_t)n * Do * Ho * Wo * ldgy + c;
        float acc = 0.f;
        for (int a = 0; a < 7; ++a) {
            if (wd[a] == 0.f) continue;
            float accd = 0.f;
            for (int b = 0; b < 7; ++b) {
                if (wh[b] == 0.f) continue;
                float acch = 0.f;
                for (int e = 0; e < 7; ++e) {
                    if (ww[e] == 0.f) continue;
                    acch = fmaf(ww[e], gb[(((int64_t)(2 * d - 3 + a) * Ho + (2 * h - 3 + b)) * Wo + (2 * w - 3 + e)) * ldgy], acch);
                }
                accd = fmaf(wh[b], acch, accd);
            }
            acc = fmaf(wd[a], accd, acc);
        }
        gx[(i / C) * ldgx + c] = acc;
    }
}

// ---- 128-bit variants (C, ld multiples of 4, 16-byte aligned bases): one thread = one voxel x four channels.  The per-channel
// arithmetic (order of comparisons / association of the sums) is exactly that of the scalar kernels above: bit-identical.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__global__ void __launch_bounds__(256)
maxpool2_fwd_v4_kernel(const float* __restrict__ x, int N, int D, int H, int W, int C, int ldx, float* __restrict__ y, int ldy) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2, C4 = C / 4;
    const int64_t total = (int64_t)N * Do * Ho * Wo * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        int64_t t = i / C4;
        const int ow = (int)(t % Wo); t /= Wo;
        const int oh = (int)(t % Ho); t /= Ho;
        const int od = (int)(t % Do);
        const int n = (int)(t / Do);
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float4 v4 = sp_ldg_stream(reinterpret_cast<const float4*>(
                        x + ((((int64_t)n * D + 2 * od + a) * H + 2 * oh + b) * W + 2 * ow + e) * ldx + c));
                    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (v[k] > m[k] || v[k] != v[k]) m[k] = v[k];
                }
        st4(y + (i / C4) * ldy + c, make_float4(m[0], m[1], m[2], m[3]));
    }
}

__global__ void __launch_bounds__(256)
maxpool2_bwd_v4_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy, int N, int D,
                       int H, int W, int C, float* __restrict__ gx) {
    const int Do = D / 2, Ho = H / 2, Wo = W / 2, C4 = C / 4;
    const int64_t total = (int64_t)N * D * H * W * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        int64_t t = i / C4;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H); t /= H;
        const int d = (int)(t % D);
        const int n = (int)(t / D);
        const int od = d >> 1, oh = h >> 1, ow = w >> 1;
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (od < Do && oh < Ho && ow < Wo) {
            const int64_t o = ((((int64_t)n * Do + od) * Ho + oh) * Wo + ow) * C + c;
            const float4 m4 = ld4(y + o), v4 = ld4(x + (i / C4) * C + c);
            const float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
            bool cand[4];
            bool any = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) { cand[k] = (v[k] == m[k]); any |= cand[k]; }
            if (any) {
                const int me = ((d & 1) << 2) | ((h & 1) << 1) | (w & 1);
                for (int q = 0; q < me; ++q) {          // an earlier element of the window already holds the maximum?
                    const int a = q >> 2, b = (q >> 1) & 1, e = q & 1;
                    const float4 u4 = ld4(x + ((((int64_t)n * D + 2 * od + a) * H + 2 * oh + b) * W + 2 * ow + e) * C + c);
                    const float u[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (u[k] == m[k]) cand[k] = false;
                }
                const float4 g4 = ld4(gy + o);
                const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (cand[k]) r[k] = g[k];
            }
        }
        st4(gx + (i / C4) * C + c, make_float4(r[0], r[1], r[2], r[3]));
    }
}

__global__ void __launch_bounds__(256)
upsample2_fwd_v4_kernel(const float* __restrict__ x, int N, int D, int H, int W, int C, int ldx, float* __restrict__ y,
                        int ldy, int align) {
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W, C4 = C / 4;
    const int64_t total = (int64_t)N * Do * Ho * Wo * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        int64_t t = i / C4;
        const int ow = (int)(t % Wo); t /= Wo;
        const int oh = (int)(t % Ho); t /= Ho;
        const int od = (int)(t % Do);
        const int n = (int)(t / Do);
        int d0, d1, h0, h1, w0, w1;
        float ld0, ld1, lh0, lh1, lw0, lw1;
        up_src(od, D, Do, align, d0, d1, ld0, ld1);
        up_src(oh, H, Ho, align, h0, h1, lh0, lh1);
        up_src(ow, W, Wo, align, w0, w1, lw0, lw1);
        const float* xb = x + (int64_t)n * D * H * W * ldx + c;
#define XV4(dd, hh, ww) ld4(xb + (((int64_t)(dd) * H + (hh)) * W + (ww)) * ldx)
        const float4 a000 = XV4(d0, h0, w0), a001 = XV4(d0, h0, w1), a010 = XV4(d0, h1, w0), a011 = XV4(d0, h1, w1);
        const float4 a100 = XV4(d1, h0, w0), a101 = XV4(d1, h0, w1), a110 = XV4(d1, h1, w0), a111 = XV4(d1, h1, w1);
#undef XV4
#define TRI(f) (ld0 * (lh0 * (lw0 * a000.f + lw1 * a001.f) + lh1 * (lw0 * a010.f + lw1 * a011.f)) + \
                ld1 * (lh0 * (lw0 * a100.f + lw1 * a101.f) + lh1 * (lw0 * a110.f + lw1 * a111.f)))
        st4(y + (i / C4) * ldy + c, make_float4(TRI(x), TRI(y), TRI(z), TRI(w)));
#undef TRI
    }
}

__global__ void __launch_bounds__(256)
upsample2_bwd_v4_kernel(const float* __restrict__ gy, int ldgy, int N, int D, int H, int W, int C, float* __restrict__ gx,
                        int ldgx, int align) {
    const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W, C4 = C / 4;
    const int64_t total = (int64_t)N * D * H * W * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        int64_t t = i / C4;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H); t /= H;
        const int d = (int)(t % D);
        const int n = (int)(t / D);
        float wd[7], wh[7], ww[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            int i0, i1;
            float l0, l1;
            int o = 2 * d - 3 + q;
            wd[q] = 0.f;
            if (o >= 0 && o < Do) {
                up_src(o, D, Do, align, i0, i1, l0, l1);
                wd[q] = (i0 == d ? l0 : 0.f) + (i1 == d ? l1 : 0.f);
            }
            o = 2 * h - 3 + q;
            wh[q] = 0.f;
            if (o >= 0 && o < Ho) {
                up_src(o, H, Ho, align, i0, i1, l0, l1);
                wh[q] = (i0 == h ? l0 : 0.f) + (i1 == h ? l1 : 0.f);
            }
            o = 2 * w - 3 + q;
            ww[q] = 0.f;
            if (o >= 0 && o < Wo) {
                up_src(o, W, Wo, align, i0, i1, l0, l1);
                ww[q] = (i0 == w ? l0 : 0.f) + (i1 == w ? l1 : 0.f);
            }
        }
        const float* gb = gy + (int64_t)n * Do * Ho * Wo * ldgy + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int a = 0; a < 7; ++a) {
            if (wd[a] == 0.f) continue;
            float4 accd = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int b = 0; b < 7; ++b) {
                if (wh[b] == 0.f) continue;
                float4 acch = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int e = 0; e < 7; ++e) {
                    if (ww[e] == 0.f) continue;
                    const float4 g = ld4(gb + (((int64_t)(2 * d - 3 + a) * Ho + (2 * h - 3 + b)) * Wo + (2 * w - 3 + e)) * ldgy);
                    acch.x = fmaf(ww[e], g.x, acch.x); acch.y = fmaf(ww[e], g.y, acch.y);
                    acch.z = fmaf(ww[e], g.z, acch.z); acch.w = fmaf(ww[e], g.w, acch.w);
                }
                accd.x = fmaf(wh[b], acch.x, accd.x); accd.y = fmaf(wh[b], acch.y, accd.y);
                accd.z = fmaf(wh[b], acch.z, accd.z); accd.w = fmaf(wh[b], acch.w, accd.w);
            }
            acc.x = fmaf(wd[a], accd.x, acc.x); acc.y = fmaf(wd[a], accd.y, acc.y);
            acc.z = fmaf(wd[a], accd.z, acc.z); acc.w = fmaf(wd[a], accd.w, acc.w);
        }
        st4(gx + (i / C4) * ldgx + c, acc);
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- crop copy / add ---------------------------------------------------------------------------------------------
template <bool ADD_INTO_BIG>
__global__ void __launch_bounds__(256)
crop_kernel(float* __restrict__ big, int Db, int Hb, int Wb, int ldb, float* __restrict__ small, int Ds, int Hs, int Ws,
            int lds, int N, int C, int od, int oh, int ow) {
    const int64_t total = (int64_t)N * Ds * Hs * Ws * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int w = (int)(t % Ws); t /= Ws;
        const int h = (int)(t % Hs); t /= Hs;
        const int d = (int)(t % Ds);
        const int n = (int)(t / Ds);
        const int64_t bi = ((((int64_t)n * Db + d + od) * Hb + h + oh) * Wb + w + ow) * ldb + c;
        const int64_t si = (i / C) * lds + c;
        if (ADD_INTO_BIG) big[bi] += small[si];
        else small[si] = big[bi];
    }
}

// ---- layout ------------------------------------------------------------------------------------------------------
// 32x32 shared-memory transpose of the [rows][cols] matrix of every sample (coalesced on both sides).  Tiles are
// linearised on grid.x so volumes of any size fit.
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int64_t cols, int64_t tiles_r,
                 int64_t tiles_c, int64_t total_tiles) {
    // src: [n][rows][cols] -> dst: [n][cols][rows]
    __shared__ float tile[32][33];
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
    for (int64_t tI = blockIdx.x; tI < total_tiles; tI += gridDim.x) {
        const int64_t tc = tI % tiles_c;
        const int64_t tr = (tI / tiles_c) % tiles_r;
        const int64_t n = tI / (tiles_c * tiles_r);
        const int64_t r0 = tr * 32, c0 = tc * 32;
        const float* s = src + n * rows * cols;
        float* d = dst + n * rows * cols;
        for (int j = ty; j < 32; j += 8) {
            const int64_t r = r0 + j, c = c0 + tx;
            if (r < rows && c < cols) tile[j][tx] = s[r * cols + c];
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int64_t c = c0 + j, r = r0 + tx;
            if (r < rows && c < cols) d[c * rows + r] = tile[tx][j];
        }
        __syncthreads();
    }
}

int launch_transpose(const float* src, float* dst, int N, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows == 1 || cols == 1) {
        SP_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * N * rows * cols, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const int64_t tr = sp_cdiv(rows, 32), tc = sp_cdiv(cols, 32);
    const int64_t total = tr * tc * N;
    int64_t grid = total;
    const int64_t cap = (int64_t)sp_num_sms() * 64;
    if (grid > cap) grid = cap;
    transpose_kernel<<<(int)grid, 256, 0, st>>>(src, dst, rows, cols, tr, tc, total);
    SP_LAUNCH_OK("transpose_kernel");
    return 0;
}

}  // namespace

extern "C" {

int sp_maxpool2_fwd(const float* x, int N, int D, int H, int W, int C, int ldx, float* y, int ldy, void* stream) {
    SP_REQUIRE(x && y, "sp_maxpool2_fwd: NULL pointer");
    SP_REQUIRE(N > 0 && D >= 2 && H >= 2 && W >= 2 && C > 0 && ldx >= C && ldy >= C, "sp_maxpool2_fwd: bad shape");
    const int64_t total = (int64_t)N * (D / 2) * (H / 2) * (W / 2) * C;
    if (C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16(x) && al16(y))
        maxpool2_fwd_v4_kernel<<<ew_grid(total / 4), 256, 0, sp_stream(stream)>>>(x, N, D, H, W, C, ldx, y, ldy);
    else
        maxpool2_fwd_kernel<<<ew_grid(total), 256, 0, sp_stream(stream)>>>(x, N, D, H, W, C, ldx, y, ldy);
    SP_LAUNCH_OK("maxpool2_fwd_kernel");
    return 0;
}

int sp_maxpool2_bwd(const float* x, const float* y, const float* gy, int N, int D, int H, int W, int C, float* gx, void* stream) {
    SP_REQUIRE(x && y && gy && gx, "sp_maxpool2_bwd: NULL pointer");
    SP_REQUIRE(N > 0 && D >= 2 && H >= 2 && W >= 2 && C > 0, "sp_maxpool2_bwd: bad shape");
    const int64_t total = (int64_t)N * D * H * W * C;
    if (C % 4 == 0 && al16(x) && al16(y) && al16(gy) && al16(gx))
        maxpool2_bwd_v4_kernel<<<ew_grid(total / 4), 256, 0, sp_stream(stream)>>>(x, y, gy, N, D, H, W, C, gx);
    else
        maxpool2_bwd_kernel<<<ew_grid(total), 256, 0, sp_stream(stream)>>>(x, y, gy, N, D, H, W, C, gx);
    SP_LAUNCH_OK("maxpool2_bwd_kernel");
    return 0;
}

int sp_upsample2_fwd(const float* x, int N, int D, int H, int W, int C, int ldx, float* y, int ldy, int align_corners, void* stream) {
    SP_REQUIRE(x && y, "sp_upsample2_fwd: NULL pointer");
    SP_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && ldx >= C && ldy >= C, "sp_upsample2_fwd: bad shape");
    const int64_t total = (int64_t)N * D * H * W * 8 * C;
    if (C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16(x) && al16(y))
        upsample2_fwd_v4_kernel<<<ew_grid(total / 4), 256, 0, sp_stream(stream)>>>(x, N, D, H, W, C, ldx, y, ldy, align_corners);
    else
        upsample2_fwd_kernel<<<ew_grid(total), 256, 0, sp_stream(stream)>>>(x, N, D, H, W, C, ldx, y, ldy, align_corners);
    SP_LAUNCH_OK("upsample2_fwd_kernel");
    return 0;
}

int sp_upsample2_bwd(const float* gy, int ldgy, int N, int D, int H, int W, int C, float* gx, int ldgx, int align_corners, void* stream) {
    SP_REQUIRE(gy && gx, "sp_upsample2_bwd: NULL pointer");
    SP_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && ldgy >= C && ldgx >= C, "sp_upsample2_bwd: bad shape");
    const int64_t total = (int64_t)N * D * H * W * C;
    if (C % 4 == 0 && ldgy % 4 == 0 && ldgx % 4 == 0 && al16(gy) && al16(gx))
        upsample2_bwd_v4_kernel<<<ew_grid(total / 4), 256, 0, sp_stream(stream)>>>(gy, ldgy, N, D, H, W, C, gx, ldgx, align_corners);
    else
        upsample2_bwd_kernel<<<ew_grid(total), 256, 0, sp_stream(stream)>>>(gy, ldgy, N, D, H, W, C, gx, ldgx, align_corners);
    SP_LAUNCH_OK("upsample2_bwd_kernel");
    return 0;
}

int sp_crop_copy(const float* src, int Ds, int Hs, int Ws, int lds, float* dst, int Dd, int Hd, int Wd, int ldd, int N, int C,
                 int od, int oh, int ow, void* stream) {
    SP_REQUIRE(src && dst, "sp_crop_copy: NULL pointer");
    SP_REQUIRE(od >= 0 && oh >= 0 && ow >= 0 && od + Dd <= Ds && oh + Hd <= Hs && ow + Wd <= Ws, "sp_crop_copy: crop window outside source");
    SP_REQUIRE(N > 0 && C > 0 && lds >= C && ldd >= C, "sp_crop_copy: bad shape");
    const int64_t total = (int64_t)N * Dd * Hd * Wd * C;
    crop_kernel<false><<<ew_grid(total), 256, 0, sp_stream(stream)>>>(const_cast<float*>(src), Ds, Hs, Ws, lds, dst, Dd, Hd, Wd, ldd,
                                                                     N, C, od, oh, ow);
    SP_LAUNCH_OK("crop_kernel<copy>");
    return 0;
}

int sp_crop_add(float* big, int Db, int Hb, int Wb, int ldb, const float* small, int Ds, int Hs, int Ws, int lds, int N, int C,
                int od, int oh, int ow, void* stream) {
    SP_REQUIRE(big && small, "sp_crop_add: NULL pointer");
    SP_REQUIRE(od >= 0 && oh >= 0 && ow >= 0 && od + Ds <= Db && oh + Hs <= Hb && ow + Ws <= Wb, "sp_crop_add: window outside target");
    SP_REQUIRE(N > 0 && C > 0 && lds >= C && ldb >= C, "sp_crop_add: bad shape");
    const int64_t total = (int64_t)N * Ds * Hs * Ws * C;
    crop_kernel<true><<<ew_grid(total), 256, 0, sp_stream(stream)>>>(big, Db, Hb, Wb, ldb, const_cast<float*>(small), Ds, Hs, Ws, lds,
                                                                    N, C, od, oh, ow);
    SP_LAUNCH_OK("crop_kernel<add>");
    return 0;
}

int sp_ncdhw_to_ndhwc(const float* src, float* dst, int N, int C, int64_t vox, void* stream) {
    SP_REQUIRE(src && dst && N > 0 && C > 0 && vox > 0, "sp_ncdhw_to_ndhwc: bad arguments");
    return launch_transpose(src, dst, N, C, vox, sp_stream(stream));
}

int sp_ndhwc_to_ncdhw(const float* src, float* dst, int N, int C, int64_t vox, void* stream) {
    SP_REQUIRE(src && dst && N > 0 && C > 0 && vox > 0, "sp_ndhwc_to_ncdhw: bad arguments");
    return launch_transpose(src, dst, N, vox, C, sp_stream(stream));   // [vox][C] -> [C][vox]
}

}  // extern "C"
