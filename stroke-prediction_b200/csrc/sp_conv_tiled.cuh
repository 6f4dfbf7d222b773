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

constexpr int TW = 16, TH = 8;                   // output tile in w, h; depth TD (4 or 8) is a template parameter
constexpr int IW = TW + 2, IH = TH + 2;
constexpr int RW = 19;                           // padded row length (float4 units), odd -> conflict-free row stride
constexpr int COT = 16;                          // output channels per thread / per CTA pass
constexpr int VT = 4;                            // voxels per thread along w

// float4 per channel-quad plane; (TD+2)*190 is 1900 (TD=8) or 1140 (TD=4), both == 4 mod 8 -> conflict-free staging
template <int TD>
__host__ __device__ constexpr int plane_f4() { return (TD + 2) * IH * RW; }

template <int CK, int TD>
constexpr size_t smem_bytes() { return (size_t)(CK / 4) * plane_f4<TD>() * 16 + (size_t)27 * CK * COT * 4; }

// d: correlation geometry (k = 3, s = 1).  wp: packed [tap][src channel][dstP]; flip != 0 reads tap 26 - t.
template <int CK, int TD>
__global__ void __launch_bounds__(32 * TD, 512 / (32 * TD))
corr3_tiled_kernel(SpConvDesc d, int nPerG, int dstP, int tiles_w, int tiles_h, int tiles_d, const float* __restrict__ src,
                   const float* __restrict__ wp, int flip, const float* __restrict__ bias, const float* __restrict__ scale,
                   const float* __restrict__ shift, float* __restrict__ dst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int ID = TD + 2;
    constexpr int PLANE = plane_f4<TD>();
    constexpr int NT = 32 * TD;
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

    float2 acc[VT][COT / 2];
#pragma unroll
    for (int v = 0; v < VT; ++v)
#pragma unroll
        for (int j = 0; j < COT / 2; ++j) acc[v][j] = make_float2(0.f, 0.f);

    const bool vec = (d.Ci % 4 == 0) && (d.ldi % 4 == 0);
    const int id0 = od0 - d.pd, ih0 = oh0 - d.ph, iw0 = ow0 - d.pw;
    const float* srcn = src + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;

    for (int c0 = 0; c0 < d.Ci; c0 += CK) {
        __syncthreads();   // previous chunk fully consumed
        // ---- stage the input halo tile of channels [c0, c0 + CK): BN applied, padding written as zeros
        constexpr int NQ = CK / 4;
        for (int i = threadIdx.x; i < ID * IH * IW * NQ; i += NT) {
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
        for (int i = threadIdx.x; i < 27 * CK * (COT / 4); i += NT) {
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
                            float2 w[COT / 2];
#pragma unroll
                            for (int j4 = 0; j4 < COT / 4; ++j4) {
                                const float4 t4 = wv[j4];
                                w[j4 * 2 + 0] = make_float2(t4.x, t4.y);
                                w[j4 * 2 + 1] = make_float2(t4.z, t4.w);
                            }
#pragma unroll
                            for (int v = 0; v < VT; ++v) {
                                const float4 xv = xin[v + kw];
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
                o.x = sp_act_fwd(acc[v][j4 * 2 + 0].x + b[j4 * 4 + 0], d.act, d.alpha);
                o.y = sp_act_fwd(acc[v][j4 * 2 + 0].y + b[j4 * 4 + 1], d.act, d.alpha);
                o.z = sp_act_fwd(acc[v][j4 * 2 + 1].x + b[j4 * 4 + 2], d.act, d.alpha);
                o.w = sp_act_fwd(acc[v][j4 * 2 + 1].y + b[j4 * 4 + 3], d.act, d.alpha);
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

template <int CK, int TD>
static inline int sp_tiled_corr_launch_t(const SpConvDesc* d, int nPerG, const float* src, const float* wp, int flip,
                                         const float* bias, const float* scale, const float* shift, float* dst,
                                         cudaStream_t st) {
    using namespace sp_tiled;
    const int tiles_w = (d->Wo + TW - 1) / TW, tiles_h = (d->Ho + TH - 1) / TH, tiles_d = (d->Do + TD - 1) / TD;
    const int64_t nblk = (int64_t)tiles_w * tiles_h * tiles_d * d->N;
    SP_REQUIRE(nblk < (1LL << 31), "tiled corr: too many tiles");
    const int dstP = (d->Co + 15) / 16 * 16;
    dim3 grid((unsigned)nblk, (unsigned)(dstP / COT));
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(corr3_tiled_kernel<CK, TD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes<CK, TD>()));
        attr = true;
    }
    corr3_tiled_kernel<CK, TD><<<grid, 32 * TD, smem_bytes<CK, TD>(), st>>>(*d, nPerG, dstP, tiles_w, tiles_h, tiles_d, src, wp,
                                                                          flip, bias, scale, shift, dst);
    SP_LAUNCH_OK("corr3_tiled_kernel");
    return 0;
}

static inline int sp_tiled_corr_launch(const SpConvDesc* d, int nPerG, const float* src, const float* wp, int flip,
                                       const float* bias, const float* scale, const float* shift, float* dst,
                                       cudaStream_t st) {
    // depth tile: 4 planes (128-thread CTAs, 4 per SM) when that wastes fewer padded planes than 8 (D = 28: 0 vs 4)
    const bool td4 = ((d->Do + 3) / 4 * 4) < ((d->Do + 7) / 8 * 8);
    if (d->Ci > 4) {
        return td4 ? sp_tiled_corr_launch_t<8, 4>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st)
                   : sp_tiled_corr_launch_t<8, 8>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st);
    }
    return td4 ? sp_tiled_corr_launch_t<4, 4>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st)
               : sp_tiled_corr_launch_t<4, 8>(d, nPerG, src, wp, flip, bias, scale, shift, dst, st);
}

// =====================================================================================================================
// Tiled weight gradient for the same 3x3x3 stride-1 layers:
//   dW[co][ci][tap] = sum_{n,o} gz[n,o,co] * xbn[n, o - p + tap, ci]
// A CTA owns one (input-channel chunk of CK, output-channel pass of 16) slab of dW and walks over output tiles of
// 16(w) x 8(h) x 4(d) voxels (grid-strided, so the register accumulators live across many tiles and only
// gridDim.x partial slabs reach memory).  Per tile it stages the BN-applied, zero-padded input halo (18 x 10 x 6) as
// channel-quad planes and the gz tile as [voxel][16].  Thread = (sub-group sg = depth plane 0..3, item = (tap, quad)):
// a 4(ci) x 16(co) register tile, fed per voxel by one LDS.128 of x (tap-shifted) and four warp-broadcast LDS.128 of
// gz: 64 FFMA per 5 LDS.  The four sub-groups are summed through shared memory at the end (fixed order), partial
// slabs are reduced across CTAs by wgrad_reduce_kernel in fixed order: deterministic, no atomics.
namespace sp_tiled {

constexpr int WTD = 4;                               // output tile depth for wgrad
constexpr int WID = WTD + 2;
constexpr int WPLANE = WID * IH * RW;                // 1140 float4 per channel-quad plane
constexpr int WVOX = TW * TH * WTD;                  // 512 output voxels per tile

template <int CK>
constexpr size_t wgrad_smem_bytes() {
    size_t stage = (size_t)(CK / 4) * WPLANE * 16 + (size_t)WVOX * COT * 4;
    size_t red = (size_t)4 * 27 * (CK / 4) * 64 * 4;
    return stage > red ? stage : red;
}

template <int CK>
__global__ void __launch_bounds__(256, 2)
wgrad3_tiled_kernel(SpConvDesc d, int nPerG, int tiles_w, int tiles_h, int tiles_d, int total_tiles, int n_co_pass,
                    const float* __restrict__ iside, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ oside, float* __restrict__ ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NQ = CK / 4;
    constexpr int NITEM = 27 * NQ;
    float4* xs = reinterpret_cast<float4*>(smem_raw);                                  // [NQ][WID][IH][RW]
    float4* gs = reinterpret_cast<float4*>(smem_raw + (size_t)NQ * WPLANE * 16);       // [WVOX][4]

    const int c0 = (blockIdx.y / n_co_pass) * CK;
    const int co0 = (blockIdx.y % n_co_pass) * COT;
    const int sg = threadIdx.x >> 6;            // depth plane of the tile handled by this 64-thread sub-group
    const int item = threadIdx.x & 63;
    const bool active = item < NITEM;
    const int tap = active ? item / NQ : 0;
    const int q = active ? item % NQ : 0;
    const int kw = tap % 3, kh = (tap / 3) % 3, kd = tap / 9;

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
        const int id0 = od0 - d.pd, ih0 = oh0 - d.ph, iw0 = ow0 - d.pw;
        const int g = n / nPerG;
        const float* srcn = iside + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;
        const float* gzn = oside + (int64_t)n * d.Do * d.Ho * d.Wo * d.ldo;
        __syncthreads();   // previous tile fully consumed
        for (int i = threadIdx.x; i < WID * IH * IW * NQ; i += 256) {
            const int qq = i % NQ;
            int r = i / NQ;
            const int iw = r % IW; r /= IW;
            const int ih = r % IH;
            const int idd = r / IH;
            const int gd = id0 + idd, gh = ih0 + ih, gw = iw0 + iw;
            const int c = c0 + qq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi && c < d.Ci) {
                const float* p = srcn + (((int64_t)gd * d.Hi + gh) * d.Wi + gw) * d.ldi + c;
                if (vec_i) {
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
            xs[qq * WPLANE + (idd * IH + ih) * RW + iw] = v;
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
                const float* p = gzn + (((int64_t)od * d.Ho + oh) * d.Wo + ow) * d.ldo + co0 + j4 * 4;
                if (vec_o) {
                    v = *reinterpret_cast<const float4*>(p);
                } else {
                    float e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = (co0 + j4 * 4 + u < d.Co) ? p[u] : 0.f;
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
            gs[i] = v;    // i == ((dd*TH + h)*TW + w)*4 + j4
        }
        __syncthreads();
        if (active) {
            const float4* xrow = xs + q * WPLANE + ((sg + kd) * IH + kh) * RW + kw;
            const float4* grow = gs + (sg * TH * TW) * (COT / 4);
#pragma unroll 1
            for (int h = 0; h < TH; ++h) {
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

    // ---- reduce the four sub-groups through shared memory (fixed order) and emit this CTA's partial slab
    __syncthreads();
    float* red = reinterpret_cast<float*>(smem_raw);     // [4][NITEM][4][COT]
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
    const int64_t wn = (int64_t)d.Co * d.Ci * 27;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = threadIdx.x; i < NITEM * 64; i += 256) {
        const int b = i % COT;
        const int a = (i / COT) % 4;
        const int it = i / 64;
        const int tp = it / NQ, qq = it % NQ;
        const int ci = c0 + qq * 4 + a, co = co0 + b;
        if (ci < d.Ci && co < d.Co) {
            const float v = ((red[(0 * NITEM + it) * 64 + a * COT + b] + red[(1 * NITEM + it) * 64 + a * COT + b]) +
                             red[(2 * NITEM + it) * 64 + a * COT + b]) + red[(3 * NITEM + it) * 64 + a * COT + b];
            wsp[((int64_t)co * d.Ci + ci) * 27 + tp] = v;
        }
    }
}

struct WgradTiledPlan {
    int ck, n_chunks, n_co_pass, tiles_w, tiles_h, tiles_d, total_tiles, grid_x;
};

static inline WgradTiledPlan wgrad_tiled_plan(const SpConvDesc* d) {
    WgradTiledPlan p;
    p.ck = d->Ci > 4 ? 8 : 4;
    p.n_chunks = (d->Ci + p.ck - 1) / p.ck;
    p.n_co_pass = (d->Co + COT - 1) / COT;
    p.tiles_w = (d->Wo + TW - 1) / TW;
    p.tiles_h = (d->Ho + TH - 1) / TH;
    p.tiles_d = (d->Do + WTD - 1) / WTD;
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_d * d->N;
    int gx = (2 * 148) / (p.n_chunks * p.n_co_pass);     // ~2 resident CTAs per SM over all slabs
    if (gx < 1) gx = 1;
    if (gx > p.total_tiles) gx = p.total_tiles;
    p.grid_x = gx;
    return p;
}

}  // namespace sp_tiled

__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int chunks, int64_t wn, float* __restrict__ dw, float beta);

static inline bool sp_tiled_wgrad_supported(const SpConvDesc* d) {
    if (d->k != 3 || d->s != 1 || sp_tiled_disabled()) return false;
    const int64_t ov = (int64_t)d->Do * d->Ho * d->Wo;
    return ov >= 2048 && d->Wo >= 8 && d->Ho >= 4;
}

static inline size_t sp_tiled_wgrad_workspace_bytes(const SpConvDesc* d) {
    if (!sp_tiled_wgrad_supported(d)) return 0;
    const sp_tiled::WgradTiledPlan p = sp_tiled::wgrad_tiled_plan(d);
    return (size_t)p.grid_x * d->Co * d->Ci * 27 * sizeof(float);
}

static inline int sp_tiled_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* scale,
                                        const float* shift, const float* oside, float* dw, float beta, float* ws,
                                        cudaStream_t st) {
    using namespace sp_tiled;
    const WgradTiledPlan p = wgrad_tiled_plan(d);
    dim3 grid(p.grid_x, p.n_chunks * p.n_co_pass);
    static bool attr8 = false, attr4 = false;
    if (p.ck == 8) {
        if (!attr8) {
            SP_CUDA(cudaFuncSetAttribute(wgrad3_tiled_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem_bytes<8>()));
            attr8 = true;
        }
        wgrad3_tiled_kernel<8><<<grid, 256, wgrad_smem_bytes<8>(), st>>>(*d, nPerG, p.tiles_w, p.tiles_h, p.tiles_d, p.total_tiles,
                                                                        p.n_co_pass, iside, scale, shift, oside, ws);
    } else {
        if (!attr4) {
            SP_CUDA(cudaFuncSetAttribute(wgrad3_tiled_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem_bytes<4>()));
            attr4 = true;
        }
        wgrad3_tiled_kernel<4><<<grid, 256, wgrad_smem_bytes<4>(), st>>>(*d, nPerG, p.tiles_w, p.tiles_h, p.tiles_d, p.total_tiles,
                                                                        p.n_co_pass, iside, scale, shift, oside, ws);
    }
    SP_LAUNCH_OK("wgrad3_tiled_kernel");
    const int64_t wn = (int64_t)d->Co * d->Ci * 27;
    int64_t rb = (wn + 255) / 256;
    if (rb > 148 * 16) rb = 148 * 16;
    wgrad_reduce_kernel<<<(int)rb, 256, 0, st>>>(ws, p.grid_x, wn, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}
