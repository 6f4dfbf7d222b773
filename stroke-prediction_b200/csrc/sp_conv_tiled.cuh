// sp_conv_tiled.cuh — shared-memory tiled FFMA tier for the FFMA-bound correlations:
//   * K = 3, S = 1 : Cae3D.py:44,52,55,63,66,186-211 and every Block3x3x3 conv of Unet3D.py:19,22; also their stride-1
//                    dgrads / ConvTranspose3d k3 s1 (Cae3D.py:178), which sp_corrT maps here with flipped taps
//   * K = 3, S = 2 : the strided down-sampling convs Cae3D.py:48,59,70
//   * K = 2, S = 2 : the dgrad of the k2 s2 ConvTranspose3d up-sampling layers Cae3D.py:193,204
// and the weight gradients of all of them.
//
// Tile: one CTA (32 x TD threads) produces a 16(w) x 8(h) x TD(d) block of output voxels x 16 output channels.
//   warp  -> output depth plane td
//   lane  -> row th = lane & 7, column group wcol = lane >> 3 (4 groups of 4 consecutive w)
//   thread micro-tile: 4 voxels x 16 channels = 64 fp32 accumulators (32 FFMA2 registers pairs).
// The input halo tile is staged per chunk of CK input channels as CK/4 planes of float4 ("channel-quad planes"):
// BatchNorm scale/shift is applied while staging and out-of-range voxels are written as 0, so zero padding stays zero
// *after* the normalisation.  Rows are padded to an odd number of float4 so that the 8 lanes of a quarter-warp (8
// different rows) hit 8 different 16-byte bank groups: conflict-free LDS.128.
//
// Stride 2 uses a PARITY-SPLIT tile: along every axis the even input coordinates of the tile are stored first, then
// the odd ones (slot(i) = i/2 or NE + i/2).  Output o reads input 2o + t, i.e. parity t&1 at index o + t/2 — so inside
// a parity block consecutive outputs read consecutive slots and the kernel body (and its bank behaviour) is that of the
// stride-1 case with a different per-tap base offset.
#pragma once
#include <stdlib.h>
#include "sp_common.cuh"

namespace sp_tiled {

constexpr int TW = 16, TH = 8;                   // output tile in w, h; depth TD is a template parameter
constexpr int COT = 16;                          // output channels per thread / per CTA pass (wgrad; forward: template COTF)
constexpr int VT = 4;                            // voxels per thread along w

// one axis of the (parity-split) input tile of an output tile of extent T
template <int K, int S, int T>
struct Axis {
    static constexpr int IN = S * (T - 1) + K;                       // input extent
    static constexpr int NE = (S == 1) ? IN : (IN + 1) / 2;          // slots holding even coordinates (S = 2)
    __host__ __device__ static constexpr int slot(int i) { return S == 1 ? i : ((i & 1) ? NE + (i >> 1) : (i >> 1)); }
    __host__ __device__ static constexpr int tap(int t) { return slot(t); }   // base slot of tap t for output 0
};

template <int K, int S>
__host__ __device__ constexpr int row_f4() {      // padded row length (float4 units), odd
    return Axis<K, S, TW>::IN | 1;
}
// S = 1, K = 3: IN = 18 -> 19.  S = 2, K = 3: IN = 33 -> 33.  S = 2, K = 2: IN = 32 -> 33.

template <int K, int S, int TD>
__host__ __device__ constexpr int plane_f4() { return Axis<K, S, TD>::IN * Axis<K, S, TH>::IN * row_f4<K, S>(); }

template <int CK, int TD, int K, int S, int COTF>
constexpr size_t smem_bytes() { return (size_t)(CK / 4) * plane_f4<K, S, TD>() * 16 + (size_t)K * K * K * CK * COTF * 4; }

// stage one channel-quad of one input voxel: BN applied, zeros outside the volume / beyond Ci
__device__ __forceinline__ float4 stage_quad(const float* __restrict__ srcn, const SpConvDesc& d, int gd, int gh, int gw, int c,
                                             bool vec, const float* __restrict__ scale, const float* __restrict__ shift, int g) {
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
    return v;
}

// d: correlation geometry (k = K, s = S).  wp: packed [tap][src channel][dstP]; flip != 0 reads tap K^3 - 1 - t.
// COTF = output channels per CTA pass (16, 8 or 2): pass blockIdx.y covers channels [co_base + y*COTF, +COTF) — the ragged
// tail of a 24-channel layer runs as one 8-wide pass and a single-channel layer (dgrad into the 1-channel input volume,
// Cae3D.py:41) as a 2-wide pass instead of burning a full 16-wide one.
template <int CK, int TD, int K, int S, int COTF>
__global__ void __launch_bounds__(32 * TD, (S == 1 ? 512 : 256) / (32 * TD))
corr_tiled_kernel(SpConvDesc d, int nPerG, int dstP, int co_base, int tiles_w, int tiles_h, int tiles_d, const float* __restrict__ src,
                  const float* __restrict__ wp, int flip, const float* __restrict__ bias, const float* __restrict__ scale,
                  const float* __restrict__ shift, float* __restrict__ dst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using AW = Axis<K, S, TW>;
    using AH = Axis<K, S, TH>;
    using AD = Axis<K, S, TD>;
    constexpr int IW = AW::IN, IH = AH::IN, ID = AD::IN;
    constexpr int RW = row_f4<K, S>();
    constexpr int PLANE = plane_f4<K, S, TD>();
    constexpr int NT = 32 * TD;
    constexpr int K3 = K * K * K;
    constexpr int NQ = CK / 4;
    float4* xs = reinterpret_cast<float4*>(smem_raw);                                   // [NQ][ID][IH][RW]
    float* wsm = reinterpret_cast<float*>(smem_raw + (size_t)NQ * PLANE * 16);          // [K3][CK][COT]

    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th_ = t % tiles_h; t /= tiles_h;
    const int td_ = t % tiles_d;
    const int n = t / tiles_d;
    const int ow0 = tw * TW, oh0 = th_ * TH, od0 = td_ * TD;
    const int co0 = co_base + blockIdx.y * COTF;
    constexpr int COT = COTF;            // shadows the namespace constant inside this kernel
    constexpr int WV = (COT >= 4) ? 4 : 2;   // floats per weight staging / load element
    const int g = n / nPerG;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ltd = warp, lth = lane & 7, lw0 = (lane >> 3) * VT;

    float2 acc[VT][COT / 2];
#pragma unroll
    for (int v = 0; v < VT; ++v)
#pragma unroll
        for (int j = 0; j < COT / 2; ++j) acc[v][j] = make_float2(0.f, 0.f);

    const bool vec = (d.Ci % 4 == 0) && (d.ldi % 4 == 0);
    const int id0 = od0 * S - d.pd, ih0 = oh0 * S - d.ph, iw0 = ow0 * S - d.pw;
    const float* srcn = src + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;

    for (int c0 = 0; c0 < d.Ci; c0 += CK) {
        __syncthreads();   // previous chunk fully consumed
        // ---- stage the input halo tile of channels [c0, c0 + CK): BN applied, padding written as zeros
        for (int i = threadIdx.x; i < ID * IH * IW * NQ; i += NT) {
            const int q = i % NQ;
            int r = i / NQ;
            const int iw = r % IW; r /= IW;
            const int ih = r % IH;
            const int idd = r / IH;
            const float4 v = stage_quad(srcn, d, id0 + idd, ih0 + ih, iw0 + iw, c0 + q * 4, vec, scale, shift, g);
            xs[q * PLANE + (AD::slot(idd) * IH + AH::slot(ih)) * RW + AW::slot(iw)] = v;
        }
        // ---- stage the weight slab [K3][CK][COT] of this chunk / output-channel pass
        for (int i = threadIdx.x; i < K3 * CK * (COT / WV); i += NT) {
            const int j4 = i % (COT / WV);
            int r = i / (COT / WV);
            const int cl = r % CK;
            const int tap = r / CK;
            const int st = flip ? K3 - 1 - tap : tap;
            const float* wsrc = wp + ((int64_t)st * d.Ci + c0 + cl) * dstP + co0 + j4 * WV;
            if (WV == 4) {
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + cl < d.Ci) w = *reinterpret_cast<const float4*>(wsrc);
                reinterpret_cast<float4*>(wsm)[(tap * CK + cl) * (COT / 4) + j4] = w;
            } else {
                float2 w = make_float2(0.f, 0.f);
                if (c0 + cl < d.Ci) w = *reinterpret_cast<const float2*>(wsrc);
                reinterpret_cast<float2*>(wsm)[(tap * CK + cl) * (COT / 2) + j4] = w;
            }
        }
        __syncthreads();

        // ---- accumulate
#pragma unroll 1
        for (int kd = 0; kd < K; ++kd) {
#pragma unroll 1
            for (int kh = 0; kh < K; ++kh) {
                const float4* row = xs + ((ltd + AD::tap(kd)) * IH + (lth + AH::tap(kh))) * RW + lw0;
                const float* wtap = wsm + ((kd * K + kh) * K) * CK * COT;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    // S = 1: xin[j] = input lw0 + j.  S = 2: xin[j] = even slot lw0 + j (taps 0, 2), xodd[j] = odd slot (tap 1)
                    constexpr int NXE = (S == 1) ? VT + K - 1 : VT + (K - 1) / 2;
                    float4 xin[NXE];
                    float4 xodd[VT];
#pragma unroll
                    for (int j = 0; j < NXE; ++j) xin[j] = row[q * PLANE + j];
                    if (S == 2) {
#pragma unroll
                        for (int j = 0; j < VT; ++j) xodd[j] = row[q * PLANE + AW::NE + j];
                    }
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float2 w[COT / 2];
                            if (WV == 4) {
                                const float4* wv = reinterpret_cast<const float4*>(wtap + (kw * CK + q * 4 + c) * COT);
#pragma unroll
                                for (int j4 = 0; j4 < COT / 4; ++j4) {
                                    const float4 t4 = wv[j4];
                                    w[j4 * 2 + 0] = make_float2(t4.x, t4.y);
                                    w[j4 * 2 + 1] = make_float2(t4.z, t4.w);
                                }
                            } else {
                                w[0] = *reinterpret_cast<const float2*>(wtap + (kw * CK + q * 4 + c) * COT);
                            }
#pragma unroll
                            for (int v = 0; v < VT; ++v) {
                                const float4 xv = (S == 1) ? xin[v + kw] : ((kw & 1) ? xodd[v] : xin[v + (kw >> 1)]);
                                const float x = (c == 0) ? xv.x : (c == 1) ? xv.y : (c == 2) ? xv.z : xv.w;
#pragma unroll
                                for (int j = 0; j < COT / 2; ++j) acc[v][j] = sp_ffma2(x, w[j], acc[v][j]);
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
    const bool vst = (COT % 4 == 0) && (d.ldo % 4 == 0) && (co0 + COT <= d.Co);
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
                o.x = sp_act_fwd(acc[v][(j4 * 2 + 0) % (COT / 2)].x + b[(j4 * 4 + 0) % COT], d.act, d.alpha);
                o.y = sp_act_fwd(acc[v][(j4 * 2 + 0) % (COT / 2)].y + b[(j4 * 4 + 1) % COT], d.act, d.alpha);
                o.z = sp_act_fwd(acc[v][(j4 * 2 + 1) % (COT / 2)].x + b[(j4 * 4 + 2) % COT], d.act, d.alpha);
                o.w = sp_act_fwd(acc[v][(j4 * 2 + 1) % (COT / 2)].y + b[(j4 * 4 + 3) % COT], d.act, d.alpha);
                reinterpret_cast<float4*>(yp)[j4] = o;
            }
        } else {
#pragma unroll
            for (int j = 0; j < COT; ++j)
                if (co0 + j < d.Co) yp[j] = sp_act_fwd(((j & 1) ? acc[v][j >> 1].y : acc[v][j >> 1].x) + b[j], d.act, d.alpha);
        }
    }
}

}  // namespace sp_tiled

// Route a correlation to the tiled kernel when it has enough output voxels per sample to fill tiles; tiny-plane layers
// (bottleneck 3x12x12 / 1x10x10) stay on the generic kernel.
static inline bool sp_tiled_disabled() {
    static int v = -1;   // SP_DISABLE_TILED=1 forces the generic kernels (A/B parity checks of the two tiers)
    if (v < 0) {
        const char* e = getenv("SP_DISABLE_TILED");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static inline bool sp_tiled_geometry(const SpConvDesc* d) {
    return (d->k == 3 && d->s == 1) || (d->k == 3 && d->s == 2) || (d->k == 2 && d->s == 2);
}

static inline bool sp_tiled_corr_supported(const SpConvDesc* d) {
    if (!sp_tiled_geometry(d) || sp_tiled_disabled()) return false;
    const int64_t ov = (int64_t)d->Do * d->Ho * d->Wo;
    return ov >= 2048 && d->Wo >= 8 && d->Ho >= 4;
}

template <int CK, int TD, int K, int S, int COTF>
static inline int sp_tiled_corr_launch_c(const SpConvDesc* d, int nPerG, int co_base, int passes, const float* src, const float* wp,
                                         int flip, const float* bias, const float* scale, const float* shift, float* dst,
                                         cudaStream_t st) {
    using namespace sp_tiled;
    const int tiles_w = (d->Wo + TW - 1) / TW, tiles_h = (d->Ho + TH - 1) / TH, tiles_d = (d->Do + TD - 1) / TD;
    const int64_t nblk = (int64_t)tiles_w * tiles_h * tiles_d * d->N;
    SP_REQUIRE(nblk < (1LL << 31), "tiled corr: too many tiles");
    const int dstP = (d->Co + 15) / 16 * 16;
    dim3 grid((unsigned)nblk, (unsigned)passes);
    constexpr size_t smem = smem_bytes<CK, TD, K, S, COTF>();
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(corr_tiled_kernel<CK, TD, K, S, COTF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    corr_tiled_kernel<CK, TD, K, S, COTF><<<grid, 32 * TD, smem, st>>>(*d, nPerG, dstP, co_base, tiles_w, tiles_h, tiles_d, src, wp, flip,
                                                                       bias, scale, shift, dst);
    SP_LAUNCH_OK("corr_tiled_kernel");
    return 0;
}

// full 16-wide passes, then the ragged tail as one 2-, 8- or 16-wide pass
template <int CK, int TD, int K, int S>
static inline int sp_tiled_corr_launch_t(const SpConvDesc* d, int nPerG, const float* src, const float* wp, int flip,
                                         const float* bias, const float* scale, const float* shift, float* dst,
                                         cudaStream_t st) {
    // stride 2, 24 output channels (Cae3D.py:48): ONE 24-wide pass — the parity-split halo tile (the dominant cost of this
    // layer: four channel-quad chunks, each staged and synchronised separately) is staged once instead of twice (16 + 8)
    if (S == 2 && d->Co == 24 && d->ldo % 4 == 0)
        return sp_tiled_corr_launch_c<CK, TD, K, S, 24>(d, nPerG, 0, 1, src, wp, flip, bias, scale, shift, dst, st);
    const int full = d->Co / 16, rem = d->Co % 16;
    if (full > 0)
        if (int e = sp_tiled_corr_launch_c<CK, TD, K, S, 16>(d, nPerG, 0, full, src, wp, flip, bias, scale, shift, dst, st)) return e;
    if (rem == 0) return 0;
    if (rem <= 2) return sp_tiled_corr_launch_c<CK, TD, K, S, 2>(d, nPerG, full * 16, 1, src, wp, flip, bias, scale, shift, dst, st);
    if (rem <= 8) return sp_tiled_corr_launch_c<CK, TD, K, S, 8>(d, nPerG, full * 16, 1, src, wp, flip, bias, scale, shift, dst, st);
    return sp_tiled_corr_launch_c<CK, TD, K, S, 16>(d, nPerG, full * 16, 1, src, wp, flip, bias, scale, shift, dst, st);
}

static inline int sp_tiled_corr_launch(const SpConvDesc* d, int nPerG, const float* src, const float* wp, int flip,
                                       const float* bias, const float* scale, const float* shift, float* dst,
                                       cudaStream_t st) {
    if (d->s == 2) {   // parity-split tile: one channel quad per chunk (81 KB / 68 KB of shared memory), 2 CTAs per SM
        if (d->k == 3) return sp_tiled_corr_launch_t<4, 4, 3, 2>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st);
        return sp_tiled_corr_launch_t<4, 4, 2, 2>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st);
    }
    // depth tile: 4 planes (128-thread CTAs, 4 per SM) when that wastes fewer padded planes than 8 (D = 28: 0 vs 4)
    const bool td4 = ((d->Do + 3) / 4 * 4) < ((d->Do + 7) / 8 * 8);
    if (d->Ci > 4) {
        return td4 ? sp_tiled_corr_launch_t<8, 4, 3, 1>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st)
                   : sp_tiled_corr_launch_t<8, 8, 3, 1>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st);
    }
    return td4 ? sp_tiled_corr_launch_t<4, 4, 3, 1>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st)
               : sp_tiled_corr_launch_t<4, 8, 3, 1>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st);
}

// =====================================================================================================================
// Tiled weight gradient for the same layer geometries:
//   dW[co][ci][tap] = sum_{n,o} O[n,o,co] * I[n, o*S - p + tap, ci]          (O-side / I-side of SpConvDesc)
// A CTA owns one (input-channel chunk of CK, output-channel pass of 16) slab of dW and walks over output tiles of
// 16(w) x 8(h) x 4(d) voxels (grid-strided, so the register accumulators live across many tiles and only gridDim.x
// partial slabs reach memory).  Per tile it stages the (BN-applied, zero-padded, parity-split for S = 2) I-side halo as
// channel-quad planes and the O-side tile as [voxel][16].  Thread = (sub-group sg, item = (tap, quad)): a 4(ci) x 16(co)
// register tile, fed per voxel by one LDS.128 of I (tap-shifted) and four warp-broadcast LDS.128 of O: 64 FFMA per 5
// LDS.  A sub-group is 32 or 64 threads (K^3 * CK/4 items rounded up) and owns a slice of the tile's voxels; the
// sub-groups are summed through shared memory at the end (fixed order), partial slabs are reduced across CTAs by
// wgrad_reduce_kernel in fixed order: deterministic, no atomics.
namespace sp_tiled {

constexpr int WTD = 4;                               // output tile depth for wgrad
constexpr int WVOX = TW * TH * WTD;                  // 512 output voxels per tile

template <int CK, int K, int S>
struct WgradCfg {
    static constexpr int NQ = CK / 4;
    static constexpr int NITEM = K * K * K * NQ;
    static constexpr int SGT = NITEM <= 32 ? 32 : 64;         // threads per sub-group
    static constexpr int NSG = 256 / SGT;                     // 4 or 8 sub-groups
    static constexpr int HSPLIT = NSG / WTD;                  // sub-groups per depth plane (1 or 2): split of the h rows
    static constexpr int HROWS = TH / HSPLIT;
    static constexpr int PLANE = plane_f4<K, S, WTD>();
    static constexpr size_t stage_bytes = (size_t)NQ * PLANE * 16 + (size_t)WVOX * COT * 4;
    static constexpr size_t red_bytes = (size_t)NSG * NITEM * 64 * 4;
    static constexpr size_t smem = stage_bytes > red_bytes ? stage_bytes : red_bytes;
    static_assert(NITEM <= 64, "too many items per sub-group");
};

template <int CK, int K, int S>
__global__ void __launch_bounds__(256, (WgradCfg<CK, K, S>::smem > 112 * 1024) ? 1 : 2)
wgrad_tiled_kernel(SpConvDesc d, int nPerG, int tiles_w, int tiles_h, int tiles_d, int total_tiles, int n_co_pass,
                   const float* __restrict__ iside, const float* __restrict__ scale, const float* __restrict__ shift,
                   const float* __restrict__ oside, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                   float* __restrict__ ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using C = WgradCfg<CK, K, S>;
    using AW = Axis<K, S, TW>;
    using AH = Axis<K, S, TH>;
    using AD = Axis<K, S, WTD>;
    constexpr int IW = AW::IN, IH = AH::IN, ID = AD::IN;
    constexpr int RW = row_f4<K, S>();
    constexpr int NQ = C::NQ, NITEM = C::NITEM, PLANE = C::PLANE;
    constexpr int K3 = K * K * K;
    float4* xs = reinterpret_cast<float4*>(smem_raw);                                  // [NQ][ID][IH][RW]
    float4* gs = reinterpret_cast<float4*>(smem_raw + (size_t)NQ * PLANE * 16);        // [WVOX][4]

    const int c0 = (blockIdx.y / n_co_pass) * CK;
    const int co0 = (blockIdx.y % n_co_pass) * COT;
    const int sg = threadIdx.x / C::SGT;        // sub-group: depth plane sg % 4, h slice sg / 4
    const int item = threadIdx.x % C::SGT;
    const bool active = item < NITEM;
    const int tap = active ? item / NQ : 0;
    const int q = active ? item % NQ : 0;
    const int kw = tap % K, kh = (tap / K) % K, kd = tap / (K * K);
    const int plane = sg % WTD, h0 = (sg / WTD) * C::HROWS;

    float2 acc[4][COT / 2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < COT / 2; ++b) acc[a][b] = make_float2(0.f, 0.f);

    const bool vec_i = (d.Ci % 4 == 0) && (d.ldi % 4 == 0);
    const bool vec_o = (d.ldo % 4 == 0) && (co0 + COT <= d.Co);

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th_ = t % tiles_h; t /= tiles_h;
        const int td_ = t % tiles_d;
        const int n = t / tiles_d;
        const int ow0 = tw * TW, oh0 = th_ * TH, od0 = td_ * WTD;
        const int id0 = od0 * S - d.pd, ih0 = oh0 * S - d.ph, iw0 = ow0 * S - d.pw;
        const int g = n / nPerG;
        const float* srcn = iside + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;
        const float* gzn = oside + (int64_t)n * d.Do * d.Ho * d.Wo * d.ldo;
        __syncthreads();   // previous tile fully consumed
        for (int i = threadIdx.x; i < ID * IH * IW * NQ; i += 256) {
            const int qq = i % NQ;
            int r = i / NQ;
            const int iw = r % IW; r /= IW;
            const int ih = r % IH;
            const int idd = r / IH;
            const float4 v = stage_quad(srcn, d, id0 + idd, ih0 + ih, iw0 + iw, c0 + qq * 4, vec_i, scale, shift, g);
            xs[qq * PLANE + (AD::slot(idd) * IH + AH::slot(ih)) * RW + AW::slot(iw)] = v;
        }
        for (int i = threadIdx.x; i < WVOX * (COT / 4); i += 256) {
            const int j4 = i % (COT / 4);
            int r = i / (COT / 4);
            const int w = r % TW; r /= TW;
            const int h = r % TH;
            const int dd = r / TH;
            const int od = od0 + dd, oh = oh0 + h, ow = ow0 + w;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (od < d.Do && oh < d.Ho && ow < d.Wo) {
                const int cc = co0 + j4 * 4;
                const float* p = gzn + (((int64_t)od * d.Ho + oh) * d.Wo + ow) * d.ldo + cc;
                float e[4];
                if (vec_o) {
                    const float4 a = *reinterpret_cast<const float4*>(p);
                    e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = (cc + u < d.Co) ? p[u] : 0.f;
                }
                if (o_scale) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (cc + u < d.Co) e[u] = fmaf(e[u], o_scale[(int64_t)g * d.Co + cc + u], o_shift[(int64_t)g * d.Co + cc + u]);
                }
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
            gs[i] = v;    // i == ((dd*TH + h)*TW + w)*4 + j4
        }
        __syncthreads();
        if (active) {
            const float4* xrow = xs + q * PLANE + ((plane + AD::tap(kd)) * IH + (h0 + AH::tap(kh))) * RW + AW::tap(kw);
            const float4* grow = gs + ((plane * TH + h0) * TW) * (COT / 4);
#pragma unroll 1
            for (int h = 0; h < C::HROWS; ++h) {
#pragma unroll 4
                for (int w = 0; w < TW; ++w) {
                    const float4 xv = xrow[h * RW + w];
                    const float4* gp = grow + (h * TW + w) * (COT / 4);
                    float2 gzv[COT / 2];
#pragma unroll
                    for (int j4 = 0; j4 < COT / 4; ++j4) {
                        const float4 t4 = gp[j4];
                        gzv[j4 * 2 + 0] = make_float2(t4.x, t4.y);
                        gzv[j4 * 2 + 1] = make_float2(t4.z, t4.w);
                    }
#pragma unroll
                    for (int b = 0; b < COT / 2; ++b) {
                        acc[0][b] = sp_ffma2(xv.x, gzv[b], acc[0][b]);
                        acc[1][b] = sp_ffma2(xv.y, gzv[b], acc[1][b]);
                        acc[2][b] = sp_ffma2(xv.z, gzv[b], acc[2][b]);
                        acc[3][b] = sp_ffma2(xv.w, gzv[b], acc[3][b]);
                    }
                }
            }
        }
    }

    // ---- reduce the sub-groups through shared memory (fixed order) and emit this CTA's partial slab
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem_raw);     // [NSG][NITEM][4][COT]
    if (active) {
        float* r = red + ((size_t)sg * NITEM + item) * 64;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < COT / 2; ++b) {
                r[a * COT + 2 * b] = acc[a][b].x;
                r[a * COT + 2 * b + 1] = acc[a][b].y;
            }
    }
    __syncthreads();
    const int64_t wn = (int64_t)d.Co * d.Ci * K3;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = threadIdx.x; i < NITEM * 64; i += 256) {
        const int b = i % COT;
        const int a = (i / COT) % 4;
        const int it = i / 64;
        const int tp = it / NQ, qq = it % NQ;
        const int ci = c0 + qq * 4 + a, co = co0 + b;
        if (ci < d.Ci && co < d.Co) {
            float v = red[(size_t)it * 64 + a * COT + b];
#pragma unroll
            for (int s = 1; s < C::NSG; ++s) v += red[((size_t)s * NITEM + it) * 64 + a * COT + b];
            wsp[((int64_t)co * d.Ci + ci) * K3 + tp] = v;
        }
    }
}

// ---- pipelined variant -------------------------------------------------------------------------------------------------
// Same decomposition, but the tile loop is software-pipelined with cp.async: while the CTA accumulates tile i out of
// buffer b, the raw I-side halo and O-side block of tile i+1 stream into buffer b^1 (zero-filled outside the volume by the
// copy itself).  BatchNorm is applied in place after arrival (a shared-memory read-modify-write, ~30-cycle latency instead
// of a global round trip in front of every tile).  One CTA per SM owns the whole shared memory: two buffers.
// Needs 16-byte-aligned channel quads on both sides (Ci, ldi, ldo multiples of 4).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int CK, int K, int S, int WD>
struct WgradPipeCfg {
    static constexpr int NQ = CK / 4;
    static constexpr int NITEM = K * K * K * NQ;
    static constexpr int SGT = NITEM <= 32 ? 32 : (NITEM <= 64 ? 64 : 128);   // threads per sub-group
    static constexpr int NSG = 256 / SGT;
    static constexpr int ROWS = WD * TH;                        // (plane, h) rows of 16 voxels per tile
    static constexpr int RPS = ROWS / NSG;                      // rows per sub-group
    // plane stride (float4 units) == 2 (mod 4) for four quads: lanes (tap, q) of one LDS.128 phase then land in 8
    // different 16-byte bank groups (q * PLANE + {tap, tap + 1}); two quads are conflict-free with PLANE == 4 (mod 8).
    static constexpr int PLANE = plane_f4<K, S, WD>() + ((NQ == 4 && plane_f4<K, S, WD>() % 4 == 0) ? 2 : 0);
    static constexpr int VOX = TW * TH * WD;
    static constexpr size_t buf_bytes = (size_t)NQ * PLANE * 16 + (size_t)VOX * COT * 4;
    static constexpr size_t red_bytes = (size_t)NSG * NITEM * 64 * 4;
    static constexpr size_t smem = 2 * buf_bytes > red_bytes ? 2 * buf_bytes : red_bytes;
    static_assert(NITEM <= 128 && ROWS % NSG == 0, "bad wgrad pipe configuration");
};

template <int CK, int K, int S, int WD>
__global__ void __launch_bounds__(256, 1)
wgrad_tiled_pipe_kernel(SpConvDesc d, int nPerG, int tiles_w, int tiles_h, int tiles_d, int total_tiles, int n_co_pass,
                        const float* __restrict__ iside, const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ oside, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                        float* __restrict__ ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using C = WgradPipeCfg<CK, K, S, WD>;
    using AW = Axis<K, S, TW>;
    using AH = Axis<K, S, TH>;
    using AD = Axis<K, S, WD>;
    constexpr int IW = AW::IN, IH = AH::IN, ID = AD::IN;
    constexpr int RW = row_f4<K, S>();
    constexpr int NQ = C::NQ, NITEM = C::NITEM, PLANE = C::PLANE;
    constexpr int K3 = K * K * K;

    const int c0 = (blockIdx.y / n_co_pass) * CK;
    const int co0 = (blockIdx.y % n_co_pass) * COT;
    const int sg = threadIdx.x / C::SGT;
    const int item = threadIdx.x % C::SGT;
    const bool active = item < NITEM;
    const int tap = active ? item / NQ : 0;
    const int q = active ? item % NQ : 0;
    const int kw = tap % K, kh = (tap / K) % K, kd = tap / (K * K);

    float2 acc[4][COT / 2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < COT / 2; ++b) acc[a][b] = make_float2(0.f, 0.f);

    auto xs_of = [&](int b) { return reinterpret_cast<float4*>(smem_raw + (size_t)b * C::buf_bytes); };
    auto gs_of = [&](int b) { return reinterpret_cast<float4*>(smem_raw + (size_t)b * C::buf_bytes + (size_t)NQ * PLANE * 16); };

    // decode a tile index
    auto tile_origin = [&](int tile, int& n, int& od0, int& oh0, int& ow0) {
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th_ = t % tiles_h; t /= tiles_h;
        const int td_ = t % tiles_d;
        n = t / tiles_d;
        ow0 = tw * TW; oh0 = th_ * TH; od0 = td_ * WD;
    };
    // raw copies of one tile into buffer b (zero fill outside the volume / beyond the channel counts)
    auto issue = [&](int tile, int b) {
        int n, od0, oh0, ow0;
        tile_origin(tile, n, od0, oh0, ow0);
        const int id0 = od0 * S - d.pd, ih0 = oh0 * S - d.ph, iw0 = ow0 * S - d.pw;
        const float* srcn = iside + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;
        const float* gzn = oside + (int64_t)n * d.Do * d.Ho * d.Wo * d.ldo;
        float4* xs = xs_of(b);
        float4* gs = gs_of(b);
        for (int i = threadIdx.x; i < ID * IH * IW * NQ; i += 256) {
            const int qq = i % NQ;
            int r = i / NQ;
            const int iw = r % IW; r /= IW;
            const int ih = r % IH;
            const int idd = r / IH;
            const int gd = id0 + idd, gh = ih0 + ih, gw = iw0 + iw, c = c0 + qq * 4;
            const bool in = gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi && c < d.Ci;
            const float* p = in ? srcn + (((int64_t)gd * d.Hi + gh) * d.Wi + gw) * d.ldi + c : iside;
            cp_async16(&xs[qq * PLANE + (AD::slot(idd) * IH + AH::slot(ih)) * RW + AW::slot(iw)], p, in ? 16 : 0);
        }
        for (int i = threadIdx.x; i < C::VOX * (COT / 4); i += 256) {
            const int j4 = i % (COT / 4);
            int r = i / (COT / 4);
            const int w = r % TW; r /= TW;
            const int h = r % TH;
            const int dd = r / TH;
            const int od = od0 + dd, oh = oh0 + h, ow = ow0 + w, cc = co0 + j4 * 4;
            int nbytes = (od < d.Do && oh < d.Ho && ow < d.Wo) ? (d.Co - cc) * 4 : 0;
            nbytes = nbytes < 0 ? 0 : (nbytes > 16 ? 16 : nbytes);
            const float* p = nbytes ? gzn + (((int64_t)od * d.Ho + oh) * d.Wo + ow) * d.ldo + cc : oside;
            cp_async16(&gs[i], p, nbytes);
        }
        cp_async_commit();
    };
    // BatchNorm in place (padding stays zero: only in-volume voxels are touched)
    auto transform = [&](int tile, int b) {
        int n, od0, oh0, ow0;
        tile_origin(tile, n, od0, oh0, ow0);
        const int g = n / nPerG;
        if (scale) {
            const int id0 = od0 * S - d.pd, ih0 = oh0 * S - d.ph, iw0 = ow0 * S - d.pw;
            float4* xs = xs_of(b);
            for (int i = threadIdx.x; i < ID * IH * IW * NQ; i += 256) {
                const int qq = i % NQ;
                int r = i / NQ;
                const int iw = r % IW; r /= IW;
                const int ih = r % IH;
                const int idd = r / IH;
                const int gd = id0 + idd, gh = ih0 + ih, gw = iw0 + iw, c = c0 + qq * 4;
                if (gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi && c < d.Ci) {
                    float4* e = &xs[qq * PLANE + (AD::slot(idd) * IH + AH::slot(ih)) * RW + AW::slot(iw)];
                    float4 v = *e;
                    const float4 sc = *reinterpret_cast<const float4*>(scale + (int64_t)g * d.Ci + c);
                    const float4 sh = *reinterpret_cast<const float4*>(shift + (int64_t)g * d.Ci + c);
                    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                    v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                    *e = v;
                }
            }
        }
        if (o_scale) {
            float4* gs = gs_of(b);
            for (int i = threadIdx.x; i < C::VOX * (COT / 4); i += 256) {
                const int j4 = i % (COT / 4);
                int r = i / (COT / 4);
                const int w = r % TW; r /= TW;
                const int h = r % TH;
                const int dd = r / TH;
                const int cc = co0 + j4 * 4;
                if (od0 + dd < d.Do && oh0 + h < d.Ho && ow0 + w < d.Wo) {
                    float4 v = gs[i];
                    float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (cc + u < d.Co) e[u] = fmaf(e[u], o_scale[(int64_t)g * d.Co + cc + u], o_shift[(int64_t)g * d.Co + cc + u]);
                    gs[i] = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
        }
    };

    int buf = 0;
    if ((int)blockIdx.x < total_tiles) issue(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        cp_async_wait_all();
        __syncthreads();                       // buffer `buf` has landed; every thread is done with buffer buf^1
        const int next = tile + gridDim.x;
        if (next < total_tiles) issue(next, buf ^ 1);
        if (scale || o_scale) {
            transform(tile, buf);
            __syncthreads();
        }
        if (active) {
            const float4* xs = xs_of(buf);
            const float4* gs = gs_of(buf);
#pragma unroll 1
            for (int rr = 0; rr < C::RPS; ++rr) {
                const int r = sg * C::RPS + rr;
                const int plane = r / TH, h = r % TH;
                const float4* xrow = xs + q * PLANE + ((plane + AD::tap(kd)) * IH + (h + AH::tap(kh))) * RW + AW::tap(kw);
                const float4* grow = gs + ((plane * TH + h) * TW) * (COT / 4);
#pragma unroll 4
                for (int w = 0; w < TW; ++w) {
                    const float4 xv = xrow[w];
                    const float4* gp = grow + w * (COT / 4);
                    float2 gzv[COT / 2];
#pragma unroll
                    for (int j4 = 0; j4 < COT / 4; ++j4) {
                        const float4 t4 = gp[j4];
                        gzv[j4 * 2 + 0] = make_float2(t4.x, t4.y);
                        gzv[j4 * 2 + 1] = make_float2(t4.z, t4.w);
                    }
#pragma unroll
                    for (int b = 0; b < COT / 2; ++b) {
                        acc[0][b] = sp_ffma2(xv.x, gzv[b], acc[0][b]);
                        acc[1][b] = sp_ffma2(xv.y, gzv[b], acc[1][b]);
                        acc[2][b] = sp_ffma2(xv.z, gzv[b], acc[2][b]);
                        acc[3][b] = sp_ffma2(xv.w, gzv[b], acc[3][b]);
                    }
                }
            }
        }
        buf ^= 1;
    }

    // ---- reduce the sub-groups through shared memory (fixed order) and emit this CTA's partial slab
    cp_async_wait_all();
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem_raw);     // [NSG][NITEM][4][COT]
    if (active) {
        float* r = red + ((size_t)sg * NITEM + item) * 64;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < COT / 2; ++b) {
                r[a * COT + 2 * b] = acc[a][b].x;
                r[a * COT + 2 * b + 1] = acc[a][b].y;
            }
    }
    __syncthreads();
    const int64_t wn = (int64_t)d.Co * d.Ci * K3;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = threadIdx.x; i < NITEM * 64; i += 256) {
        const int b = i % COT;
        const int a = (i / COT) % 4;
        const int it = i / 64;
        const int tp = it / NQ, qq = it % NQ;
        const int ci = c0 + qq * 4 + a, co = co0 + b;
        if (ci < d.Ci && co < d.Co) {
            float v = red[(size_t)it * 64 + a * COT + b];
#pragma unroll
            for (int s2 = 1; s2 < C::NSG; ++s2) v += red[((size_t)s2 * NITEM + it) * 64 + a * COT + b];
            wsp[((int64_t)co * d.Ci + ci) * K3 + tp] = v;
        }
    }
}

struct WgradTiledPlan {
    int ck, n_chunks, n_co_pass, tiles_w, tiles_h, tiles_d, total_tiles, grid_x;
    int pipe;     // 1: wgrad_tiled_pipe_kernel (cp.async double buffering, one CTA per SM)
};

static inline bool wgrad_pipe_disabled() {
    static int v = -1;   // SP_DISABLE_WGRAD_PIPE=1 forces the non-pipelined kernel (A/B checks)
    if (v < 0) {
        const char* e = getenv("SP_DISABLE_WGRAD_PIPE");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static inline WgradTiledPlan wgrad_tiled_plan(const SpConvDesc* d) {
    WgradTiledPlan p;
    // pipelined variant: 3x3x3 layers whose channel quads are 16-byte aligned on both sides
    p.pipe = (d->k == 3 && d->Ci % 4 == 0 && d->ldi % 4 == 0 && d->ldo % 4 == 0 && !wgrad_pipe_disabled()) ? 1 : 0;
    if (p.pipe) p.ck = (d->s == 2) ? 4 : (d->Ci % 16 == 0 ? 16 : (d->Ci > 4 ? 8 : 4));
    else p.ck = (d->s == 2 && d->k == 3) ? 4 : (d->Ci > 4 ? 8 : 4);
    p.n_chunks = (d->Ci + p.ck - 1) / p.ck;
    p.n_co_pass = (d->Co + COT - 1) / COT;
    p.tiles_w = (d->Wo + TW - 1) / TW;
    p.tiles_h = (d->Ho + TH - 1) / TH;
    p.tiles_d = (d->Do + WTD - 1) / WTD;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_d * d->N;
    int gx = ((p.pipe ? 1 : 2) * 148) / (p.n_chunks * p.n_co_pass);     // resident CTAs per SM over all slabs
    if (gx < 1) gx = 1;
    if (gx > p.total_tiles) gx = p.total_tiles;
    p.grid_x = gx;
    return p;
}

}  // namespace sp_tiled

__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int chunks, int64_t wn, float* __restrict__ dw, float beta);

static inline bool sp_tiled_wgrad_supported(const SpConvDesc* d) {
    if (!sp_tiled_geometry(d) || sp_tiled_disabled()) return false;
    const int64_t ov = (int64_t)d->Do * d->Ho * d->Wo;
    return ov >= 2048 && d->Wo >= 8 && d->Ho >= 4;
}

static inline size_t sp_tiled_wgrad_workspace_bytes(const SpConvDesc* d) {
    if (!sp_tiled_wgrad_supported(d)) return 0;
    const sp_tiled::WgradTiledPlan p = sp_tiled::wgrad_tiled_plan(d);
    return (size_t)p.grid_x * d->Co * d->Ci * d->k * d->k * d->k * sizeof(float);
}

template <int CK, int K, int S>
static inline int sp_tiled_wgrad_launch_t(const SpConvDesc* d, const sp_tiled::WgradTiledPlan& p, int nPerG, const float* iside,
                                          const float* scale, const float* shift, const float* oside, const float* o_scale,
                                          const float* o_shift, float* ws, cudaStream_t st) {
    using namespace sp_tiled;
    constexpr size_t smem = WgradCfg<CK, K, S>::smem;
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(wgrad_tiled_kernel<CK, K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid(p.grid_x, p.n_chunks * p.n_co_pass);
    wgrad_tiled_kernel<CK, K, S><<<grid, 256, smem, st>>>(*d, nPerG, p.tiles_w, p.tiles_h, p.tiles_d, p.total_tiles, p.n_co_pass,
                                                         iside, scale, shift, oside, o_scale, o_shift, ws);
    SP_LAUNCH_OK("wgrad_tiled_kernel");
    return 0;
}

template <int CK, int K, int S>
static inline int sp_tiled_wgrad_pipe_launch_t(const SpConvDesc* d, const sp_tiled::WgradTiledPlan& p, int nPerG, const float* iside,
                                               const float* scale, const float* shift, const float* oside, const float* o_scale,
                                               const float* o_shift, float* ws, cudaStream_t st) {
    using namespace sp_tiled;
    constexpr size_t smem = WgradPipeCfg<CK, K, S, WTD>::smem;
    static_assert(smem <= 227 * 1024, "wgrad pipe buffers exceed shared memory");
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(wgrad_tiled_pipe_kernel<CK, K, S, WTD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid(p.grid_x, p.n_chunks * p.n_co_pass);
    wgrad_tiled_pipe_kernel<CK, K, S, WTD><<<grid, 256, smem, st>>>(*d, nPerG, p.tiles_w, p.tiles_h, p.tiles_d, p.total_tiles,
                                                                    p.n_co_pass, iside, scale, shift, oside, o_scale, o_shift, ws);
    SP_LAUNCH_OK("wgrad_tiled_pipe_kernel");
    return 0;
}

static inline int sp_tiled_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* scale,
                                        const float* shift, const float* oside, const float* o_scale, const float* o_shift,
                                        float* dw, float beta, float* ws, cudaStream_t st) {
    using namespace sp_tiled;
    const WgradTiledPlan p = wgrad_tiled_plan(d);
    int e;
    if (p.pipe && d->s == 2) e = sp_tiled_wgrad_pipe_launch_t<4, 3, 2>(d, p, nPerG, iside, scale, shift, oside, o_scale, o_shift, ws, st);
    else if (p.pipe && p.ck == 16) e = sp_tiled_wgrad_pipe_launch_t<16, 3, 1>(d, p, nPerG, iside, scale, shift, oside, o_scale, o_shift, ws, st);
    else if (p.pipe && p.ck == 8) e = sp_tiled_wgrad_pipe_launch_t<8, 3, 1>(d, p, nPerG, iside, scale, shift, oside, o_scale, o_shift, ws, st);
    else if (p.pipe) e = sp_tiled_wgrad_pipe_launch_t<4, 3, 1>(d, p, nPerG, iside, scale, shift, oside, o_scale, o_shift, ws, st);
    else if (d->s == 2 && d->k == 3) e = sp_tiled_wgrad_launch_t<4, 3, 2>(d, p, nPerG, iside, scale, shift, oside, o_scale, o_shift, ws, st);
    else if (d->s == 2 && p.ck == 8) e = sp_tiled_wgrad_launch_t<8, 2, 2>(d, p, nPerG, iside, scale, shift, oside, o_scale, o_shift, ws, st);
    else if (d->s == 2) e = sp_tiled_wgrad_launch_t<4, 2, 2>(d, p, nPerG, iside, scale, shift, oside, o_scale, o_shift, ws, st);
    else if (p.ck == 8) e = sp_tiled_wgrad_launch_t<8, 3, 1>(d, p, nPerG, iside, scale, shift, oside, o_scale, o_shift, ws, st);
    else e = sp_tiled_wgrad_launch_t<4, 3, 1>(d, p, nPerG, iside, scale, shift, oside, o_scale, o_shift, ws, st);
    if (e) return e;
    const int64_t wn = (int64_t)d->Co * d->Ci * d->k * d->k * d->k;
    int64_t rb = (wn + 255) / 256;
    if (rb > 148 * 16) rb = 148 * 16;
    wgrad_reduce_kernel<<<(int)rb, 256, 0, st>>>(ws, p.grid_x, wn, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}
