// sp_conv_tiled.cuh — shared-memory tiled fast path for the FFMA-bound 3x3x3 stride-1 correlations
// (Cae3D.py:44,52,55,63,66,186-211 and every Block3x3x3 conv of Unet3D.py:19,22; also their stride-1 dgrads, which
// sp_corrT maps onto the same kernel with flipped taps).
//
// Tile: one CTA (256 threads, 8 warps) produces a 16(w) x 8(h) x 8(d) block of output voxels x 16 output channels.
//   warp  -> output depth plane td (0..7)
//   lane  -> row th = lane & 7, column group wcol = lane >> 3 (4 groups of 4 consecutive w)
//   thread micro-tile: 4 voxels x 16 channels = 64 fp32 accumulators.
// The input halo tile (18 x 10 x 10 voxels) is staged per chunk of CK input channels as CK/4 planes of float4
// ("channel-quad planes"): BatchNorm scale/shift is applied while staging and out-of-range voxels are written as 0,
// so zero padding stays zero *after* the normalisation.  Rows are padded to 19 float4 so that the 8 lanes of a
// quarter-warp (8 different rows) hit 8 different 16-byte bank groups: conflict-free LDS.128.  The 27 x CK x 16
// weight slab of the chunk sits next to it and is read with warp-uniform (broadcast) LDS.128.
// Per (kd, kh, channel quad) a thread issues 6 + 48 LDS.128 and 768 FFMA.
#pragma once
#include <stdlib.h>
#include "sp_common.cuh"

namespace sp_tiled {

constexpr int TW = 16, TH = 8, TD = 8;           // output tile
constexpr int IW = TW + 2, IH = TH + 2, ID = TD + 2;
constexpr int RW = 19;                           // padded row length (float4 units), odd -> conflict-free row stride
constexpr int PLANE = ID * IH * RW;              // 1900 float4 per channel-quad plane; 1900 % 8 == 4 -> conflict-free staging
constexpr int COT = 16;                          // output channels per thread / per CTA pass
constexpr int VT = 4;                            // voxels per thread along w

template <int CK>
constexpr size_t smem_bytes() { return (size_t)(CK / 4) * PLANE * 16 + (size_t)27 * CK * COT * 4; }

// d: correlation geometry (k = 3, s = 1).  wp: packed [tap][src channel][dstP]; flip != 0 reads tap 26 - t.
template <int CK>
__global__ void __launch_bounds__(256, 2)
corr3_tiled_kernel(SpConvDesc d, int nPerG, int dstP, int tiles_w, int tiles_h, int tiles_d, const float* __restrict__ src,
                   const float* __restrict__ wp, int flip, const float* __restrict__ bias, const float* __restrict__ scale,
                   const float* __restrict__ shift, float* __restrict__ dst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* xs = reinterpret_cast<float4*>(smem_raw);                                   // [CK/4][ID][IH][RW]
    float* wsm = reinterpret_cast<float*>(smem_raw + (size_t)(CK / 4) * PLANE * 16);    // [27][CK][COT]

    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th_ = t % tiles_h; t /= tiles_h;
    const int td_ = t % tiles_d;
    const int n = t / tiles_d;
    const int ow0 = tw * TW, oh0 = th_ * TH, od0 = td_ * TD;
    const int co0 = blockIdx.y * COT;
    const int g = n / nPerG;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ltd = warp, lth = lane & 7, lw0 = (lane >> 3) * VT;

    float acc[VT][COT];
#pragma unroll
    for (int v = 0; v < VT; ++v)
#pragma unroll
        for (int j = 0; j < COT; ++j) acc[v][j] = 0.f;

    const bool vec = (d.Ci % 4 == 0) && (d.ldi % 4 == 0);
    const int id0 = od0 - d.pd, ih0 = oh0 - d.ph, iw0 = ow0 - d.pw;
    const float* srcn = src + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;

    for (int c0 = 0; c0 < d.Ci; c0 += CK) {
        __syncthreads();   // previous chunk fully consumed
        // ---- stage the input halo tile of channels [c0, c0 + CK): BN applied, padding written as zeros
        constexpr int NQ = CK / 4;
        for (int i = threadIdx.x; i < ID * IH * IW * NQ; i += 256) {
            const int q = i % NQ;
            int r = i / NQ;
            const int iw = r % IW; r /= IW;
            const int ih = r % IH;
            const int idd = r / IH;
            const int gd = id0 + idd, gh = ih0 + ih, gw = iw0 + iw;
            const int c = c0 + q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi && c < d.Ci) {
                const float* p = srcn + (((int64_t)gd * d.Hi + gh) * d.Wi + gw) * d.ldi + c;
                if (vec) {
                    v = *reinterpret_cast<const float4*>(p);
                    if (scale) {
                        const float4 sc = *reinterpret_cast<const float4*>(scale + (int64_t)g * d.Ci + c);
                        const float4 sh = *reinterpret_cast<const float4*>(shift + (int64_t)g * d.Ci + c);
                        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                    }
                } else {
                    float e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        e[u] = 0.f;
                        if (c + u < d.Ci) {
                            e[u] = p[u];
                            if (scale) e[u] = fmaf(e[u], scale[(int64_t)g * d.Ci + c + u], shift[(int64_t)g * d.Ci + c + u]);
                        }
                    }
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
            xs[q * PLANE + (idd * IH + ih) * RW + iw] = v;
        }
        // ---- stage the weight slab [27][CK][COT] of this chunk / output-channel pass
        for (int i = threadIdx.x; i < 27 * CK * (COT / 4); i += 256) {
            const int j4 = i % (COT / 4);
            int r = i / (COT / 4);
            const int cl = r % CK;
            const int tap = r / CK;
            const int st = flip ? 26 - tap : tap;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + cl < d.Ci) w = *reinterpret_cast<const float4*>(wp + ((int64_t)st * d.Ci + c0 + cl) * dstP + co0 + j4 * 4);
            reinterpret_cast<float4*>(wsm)[(tap * CK + cl) * (COT / 4) + j4] = w;
        }
        __syncthreads();

        // ---- accumulate
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll 1
            for (int kh = 0; kh < 3; ++kh) {
                const float4* row = xs + ((ltd + kd) * IH + (lth + kh)) * RW + lw0;
                const float* wtap = wsm + ((kd * 3 + kh) * 3) * CK * COT;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    float4 xin[VT + 2];
#pragma unroll
                    for (int j = 0; j < VT + 2; ++j) xin[j] = row[q * PLANE + j];
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4* wv = reinterpret_cast<const float4*>(wtap + (kw * CK + q * 4 + c) * COT);
                            float w[COT];
#pragma unroll
                            for (int j4 = 0; j4 < COT / 4; ++j4) {
                                const float4 t4 = wv[j4];
                                w[j4 * 4 + 0] = t4.x; w[j4 * 4 + 1] = t4.y; w[j4 * 4 + 2] = t4.z; w[j4 * 4 + 3] = t4.w;
                            }
#pragma unroll
                            for (int v = 0; v < VT; ++v) {
                                const float4 xv = xin[v + kw];
                                const float x = (c == 0) ? xv.x : (c == 1) ? xv.y : (c == 2) ? xv.z : xv.w;
#pragma unroll
                                for (int j = 0; j < COT; ++j) acc[v][j] = fmaf(x, w[j], acc[v][j]);
                            }
                        }
                    }
                }
            }
        }
    }

    // ---- epilogue: bias + activation, masked stores
    const int od = od0 + ltd, oh = oh0 + lth;
    if (od >= d.Do || oh >= d.Ho) return;
    const bool vst = (d.ldo % 4 == 0) && (co0 + COT <= d.Co);
    float b[COT];
#pragma unroll
    for (int j = 0; j < COT; ++j) b[j] = (bias && co0 + j < d.Co) ? bias[co0 + j] : 0.f;
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        const int ow = ow0 + lw0 + v;
        if (ow >= d.Wo) continue;
        float* yp = dst + ((((int64_t)n * d.Do + od) * d.Ho + oh) * d.Wo + ow) * d.ldo + co0;
        if (vst) {
#pragma unroll
            for (int j4 = 0; j4 < COT / 4; ++j4) {
                float4 o;
                o.x = sp_act_fwd(acc[v][j4 * 4 + 0] + b[j4 * 4 + 0], d.act, d.alpha);
                o.y = sp_act_fwd(acc[v][j4 * 4 + 1] + b[j4 * 4 + 1], d.act, d.alpha);
                o.z = sp_act_fwd(acc[v][j4 * 4 + 2] + b[j4 * 4 + 2], d.act, d.alpha);
                o.w = sp_act_fwd(acc[v][j4 * 4 + 3] + b[j4 * 4 + 3], d.act, d.alpha);
                reinterpret_cast<float4*>(yp)[j4] = o;
            }
        } else {
#pragma unroll
            for (int j = 0; j < COT; ++j)
                if (co0 + j < d.Co) yp[j] = sp_act_fwd(acc[v][j] + b[j], d.act, d.alpha);
        }
    }
}

}  // namespace sp_tiled

// Route a correlation to the tiled kernel when it is a 3x3x3 stride-1 layer with enough output voxels per sample to
// fill tiles; tiny-plane layers (bottleneck 3x12x12 / 1x10x10) stay on the generic kernel.
static inline bool sp_tiled_disabled() {
    static int v = -1;   // SP_DISABLE_TILED=1 forces the generic kernels (A/B parity checks of the two tiers)
    if (v < 0) {
        const char* e = getenv("SP_DISABLE_TILED");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static inline bool sp_tiled_corr_supported(const SpConvDesc* d) {
    if (d->k != 3 || d->s != 1 || sp_tiled_disabled()) return false;
    const int64_t ov = (int64_t)d->Do * d->Ho * d->Wo;
    return ov >= 2048 && d->Wo >= 8 && d->Ho >= 4;
}

static inline int sp_tiled_corr_launch(const SpConvDesc* d, int nPerG, const float* src, const float* wp, int flip,
                                       const float* bias, const float* scale, const float* shift, float* dst,
                                       cudaStream_t st) {
    using namespace sp_tiled;
    const int tiles_w = (d->Wo + TW - 1) / TW, tiles_h = (d->Ho + TH - 1) / TH, tiles_d = (d->Do + TD - 1) / TD;
    const int64_t nblk = (int64_t)tiles_w * tiles_h * tiles_d * d->N;
    SP_REQUIRE(nblk < (1LL << 31), "tiled corr: too many tiles");
    const int dstP = (d->Co + 15) / 16 * 16;
    dim3 grid((unsigned)nblk, (unsigned)(dstP / COT));
    static bool attr8 = false, attr4 = false;
    if (d->Ci > 4) {
        if (!attr8) {
            SP_CUDA(cudaFuncSetAttribute(corr3_tiled_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<8>()));
            attr8 = true;
        }
        corr3_tiled_kernel<8><<<grid, 256, smem_bytes<8>(), st>>>(*d, nPerG, dstP, tiles_w, tiles_h, tiles_d, src, wp, flip, bias,
                                                                 scale, shift, dst);
    } else {
        if (!attr4) {
            SP_CUDA(cudaFuncSetAttribute(corr3_tiled_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<4>()));
            attr4 = true;
        }
        corr3_tiled_kernel<4><<<grid, 256, smem_bytes<4>(), st>>>(*d, nPerG, dstP, tiles_w, tiles_h, tiles_d, src, wp, flip, bias,
                                                                 scale, shift, dst);
    }
    SP_LAUNCH_OK("corr3_tiled_kernel");
    return 0;
}

// ---- tiled wgrad: not enabled yet (generic two-stage kernel is used) ------------------------------------------------
static inline bool sp_tiled_wgrad_supported(const SpConvDesc*) { return false; }
static inline size_t sp_tiled_wgrad_workspace_bytes(const SpConvDesc*) { return 0; }
static inline int sp_tiled_wgrad_launch(const SpConvDesc*, int, const float*, const float*, const float*, const float*,
                                        float*, float, float*, cudaStream_t) { return -1; }
