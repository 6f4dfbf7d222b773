// sp_conv_tiledT.cuh — shared-memory tiled FFMA kernel for the STRIDE-2 transposed correlation
//   I[n,i,ci] = sum_{tap,co} O[n,(i + p - tap)/2, co] * W[co,ci,tap]      (terms with odd i + p - tap do not exist)
// i.e. the dgrad of the strided convs (Cae3D.py:48,59) and the forward of the k2 s2 ConvTranspose3d up-sampling layers
// (Cae3D.py:193,204).
//
// Which taps reach an I-side voxel depends only on the parity of its coordinates: per axis a voxel of parity class c sees
// the taps t == (c + p) mod 2 (k = 3: {0,2} or {1}; k = 2: exactly one).  A CTA owns a 32(w) x 16(h) x 8(d) I-side block =
// 16 x 8 x 4 "cells" of 2x2x2 voxels and the (16+e) x (8+e) x (4+e) O-side voxels it depends on (e = 1 for k = 3), which it
// stages once, ALL channels, as channel-quad planes (BN applied, zeros outside the volume).  Thread = (cell depth, cell
// row, group of 4 consecutive cells along w); it then walks the 8 parity classes one after the other: for each class the
// CTA stages the <= 8 taps of that class, every thread accumulates 4 voxels (same parity, so the same taps) x 16
// I-side channels in registers and stores them.  All threads are in the same class at the same time: no divergence, and
// per (tap, channel) 4 broadcast LDS.128 of weights feed 64 FFMA, like the forward tier (sp_conv_tiled.cuh).
#pragma once
#include "sp_common.cuh"

namespace sp_tiledT {

constexpr int CW = 16, CH = 8, CD = 4;      // cells per CTA (w, h, d); the I-side block is twice that per axis
constexpr int VT = 4;                       // cells per thread along w
constexpr int CIT = 16;                     // I-side channels per CTA pass
constexpr int MAXCO = 32;                   // O-side channels held in shared memory

template <int K>
struct Geo {
    static constexpr int E = (K == 3) ? 1 : 0;
    static constexpr int OW = CW + E, OH = CH + E, OD = CD + E;
    static constexpr int RW = OW | 1;                       // odd row stride (float4 units): conflict-free LDS.128
    static constexpr int PLANE = OD * OH * RW;              // float4 per channel-quad plane
};

template <int K>
static inline size_t smem_bytes(int Co) {
    const int nq = (Co + 3) / 4;
    return (size_t)nq * Geo<K>::PLANE * 16 + (size_t)8 * nq * 4 * CIT * 4;
}

template <int K>
__global__ void __launch_bounds__(32 * CD, 2)
corrT_s2_tiled_kernel(SpConvDesc d, int nPerG, int ciP, int tiles_w, int tiles_h, int tiles_d, const float* __restrict__ src,
                      const float* __restrict__ wp, const float* __restrict__ bias, const float* __restrict__ scale,
                      const float* __restrict__ shift, float* __restrict__ dst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using G = Geo<K>;
    constexpr int NT = 32 * CD;
    const int NQ = (d.Co + 3) / 4;
    const int CoP = NQ * 4;
    float4* os = reinterpret_cast<float4*>(smem_raw);                                   // [NQ][OD][OH][RW]
    float* wsm = reinterpret_cast<float*>(smem_raw + (size_t)NQ * G::PLANE * 16);       // [<= 8 taps][CoP][CIT]

    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th_ = t % tiles_h; t /= tiles_h;
    const int td_ = t % tiles_d;
    const int n = t / tiles_d;
    const int iw0 = tw * 2 * CW, ih0 = th_ * 2 * CH, id0 = td_ * 2 * CD;     // even
    const int ci0 = blockIdx.y * CIT;
    const int g = n / nPerG;
    // first O-side voxel any voxel of the block depends on: ceil((i0 + p - (K-1)) / 2)
    const int obw = (iw0 + d.pw - (K - 1) + 1) >> 1, obh = (ih0 + d.ph - (K - 1) + 1) >> 1, obd = (id0 + d.pd - (K - 1) + 1) >> 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ltd = warp, lth = lane & 7, lw0 = (lane >> 3) * VT;

    // ---- stage the O-side block, all channels
    {
        const bool vec = (d.Co % 4 == 0) && (d.ldo % 4 == 0);
        const float* srcn = src + (int64_t)n * d.Do * d.Ho * d.Wo * d.ldo;
        for (int i = threadIdx.x; i < G::OD * G::OH * G::OW * NQ; i += NT) {
            const int q = i % NQ;
            int r = i / NQ;
            const int ww = r % G::OW; r /= G::OW;
            const int hh = r % G::OH;
            const int dd = r / G::OH;
            const int od = obd + dd, oh = obh + hh, ow = obw + ww;
            const int c = q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (od >= 0 && od < d.Do && oh >= 0 && oh < d.Ho && ow >= 0 && ow < d.Wo) {
                const float* p = srcn + (((int64_t)od * d.Ho + oh) * d.Wo + ow) * d.ldo + c;
                float e[4];
                if (vec) {
                    const float4 a = *reinterpret_cast<const float4*>(p);
                    e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = (c + u < d.Co) ? p[u] : 0.f;
                }
                if (scale) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (c + u < d.Co) e[u] = fmaf(e[u], scale[(int64_t)g * d.Co + c + u], shift[(int64_t)g * d.Co + c + u]);
                }
                v = make_float4(e[0], e[1], e[2], e[3]);
            }
            os[q * G::PLANE + (dd * G::OH + hh) * G::RW + ww] = v;
        }
    }

    float b[CIT];
#pragma unroll
    for (int j = 0; j < CIT; ++j) b[j] = (bias && ci0 + j < d.Ci) ? bias[ci0 + j] : 0.f;
    const bool vst = (d.ldi % 4 == 0) && (ci0 + CIT <= d.Ci);

#pragma unroll 1
    for (int cls = 0; cls < 8; ++cls) {
        const int cw = cls & 1, chh = (cls >> 1) & 1, cd = cls >> 2;
        // taps of this class per axis: t = t0, t0 + 2 (< K); O-side offset of tap t for cell 0: (c + p - t)/2 + i0/2 - ob
        const int tw0 = (cw + d.pw) & 1, th0 = (chh + d.ph) & 1, td0 = (cd + d.pd) & 1;
        const int nw = (tw0 + 2 < K) ? 2 : 1, nh = (th0 + 2 < K) ? 2 : 1, nd = (td0 + 2 < K) ? 2 : 1;
        const int bw0 = ((cw + d.pw - tw0) >> 1) + (iw0 >> 1) - obw;       // offset for tap tw0; tap tw0 + 2 is one less
        const int bh0 = ((chh + d.ph - th0) >> 1) + (ih0 >> 1) - obh;
        const int bd0 = ((cd + d.pd - td0) >> 1) + (id0 >> 1) - obd;
        const int ntap = nw * nh * nd;

        __syncthreads();     // previous class done with wsm (and, first time, the O-side block is complete)
        for (int i = threadIdx.x; i < ntap * CoP * (CIT / 4); i += NT) {
            const int j4 = i % (CIT / 4);
            int r = i / (CIT / 4);
            const int co = r % CoP;
            const int ti = r / CoP;
            const int jw = ti % nw, jh = (ti / nw) % nh, jd = ti / (nw * nh);
            const int tap = ((td0 + 2 * jd) * K + (th0 + 2 * jh)) * K + (tw0 + 2 * jw);
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (co < d.Co) w = *reinterpret_cast<const float4*>(wp + ((int64_t)tap * d.Co + co) * ciP + ci0 + j4 * 4);
            reinterpret_cast<float4*>(wsm)[(ti * CoP + co) * (CIT / 4) + j4] = w;
        }
        __syncthreads();

        float2 acc[VT][CIT / 2];
#pragma unroll
        for (int v = 0; v < VT; ++v)
#pragma unroll
            for (int j = 0; j < CIT / 2; ++j) acc[v][j] = make_float2(0.f, 0.f);

#pragma unroll 1
        for (int jd = 0; jd < nd; ++jd) {
#pragma unroll 1
            for (int jh = 0; jh < nh; ++jh) {
#pragma unroll 1
                for (int jw = 0; jw < nw; ++jw) {
                    const float4* row = os + ((ltd + bd0 - jd) * G::OH + (lth + bh0 - jh)) * G::RW + lw0 + bw0 - jw;
                    const float* wt = wsm + (size_t)((jd * nh + jh) * nw + jw) * CoP * CIT;
#pragma unroll 1
                    for (int q = 0; q < NQ; ++q) {
                        float4 xin[VT];
#pragma unroll
                        for (int v = 0; v < VT; ++v) xin[v] = row[q * G::PLANE + v];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4* wv = reinterpret_cast<const float4*>(wt + (q * 4 + c) * CIT);
                            float2 w[CIT / 2];
#pragma unroll
                            for (int j4 = 0; j4 < CIT / 4; ++j4) {
                                const float4 t4 = wv[j4];
                                w[j4 * 2 + 0] = make_float2(t4.x, t4.y);
                                w[j4 * 2 + 1] = make_float2(t4.z, t4.w);
                            }
#pragma unroll
                            for (int v = 0; v < VT; ++v) {
                                const float x = (c == 0) ? xin[v].x : (c == 1) ? xin[v].y : (c == 2) ? xin[v].z : xin[v].w;
#pragma unroll
                                for (int j = 0; j < CIT / 2; ++j) acc[v][j] = sp_ffma2(x, w[j], acc[v][j]);
                            }
                        }
                    }
                }
            }
        }

        // ---- epilogue of this class: bias + activation, masked stores
        const int id = id0 + 2 * ltd + cd, ih = ih0 + 2 * lth + chh;
        if (id < d.Di && ih < d.Hi) {
#pragma unroll
            for (int v = 0; v < VT; ++v) {
                const int iw = iw0 + 2 * (lw0 + v) + cw;
                if (iw >= d.Wi) continue;
                float* yp = dst + ((((int64_t)n * d.Di + id) * d.Hi + ih) * d.Wi + iw) * d.ldi + ci0;
                if (vst) {
#pragma unroll
                    for (int j4 = 0; j4 < CIT / 4; ++j4) {
                        float4 o;
                        o.x = sp_act_fwd(acc[v][j4 * 2 + 0].x + b[j4 * 4 + 0], d.act, d.alpha);
                        o.y = sp_act_fwd(acc[v][j4 * 2 + 0].y + b[j4 * 4 + 1], d.act, d.alpha);
                        o.z = sp_act_fwd(acc[v][j4 * 2 + 1].x + b[j4 * 4 + 2], d.act, d.alpha);
                        o.w = sp_act_fwd(acc[v][j4 * 2 + 1].y + b[j4 * 4 + 3], d.act, d.alpha);
                        reinterpret_cast<float4*>(yp)[j4] = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CIT; ++j)
                        if (ci0 + j < d.Ci) yp[j] = sp_act_fwd(((j & 1) ? acc[v][j >> 1].y : acc[v][j >> 1].x) + b[j], d.act, d.alpha);
                }
            }
        }
    }
}

}  // namespace sp_tiledT

static inline bool sp_tiledT_supported(const SpConvDesc* d) {
    if (d->s != 2 || (d->k != 2 && d->k != 3) || sp_tiled_disabled()) return false;
    if (d->Co > sp_tiledT::MAXCO) return false;
    const int64_t iv = (int64_t)d->Di * d->Hi * d->Wi;
    return iv >= 8192 && d->Wi >= 16 && d->Hi >= 8;
}

template <int K>
static inline int sp_tiledT_launch_t(const SpConvDesc* d, int nPerG, const float* src, const float* wp, const float* bias,
                                     const float* scale, const float* shift, float* dst, cudaStream_t st) {
    using namespace sp_tiledT;
    const int tiles_w = (d->Wi + 2 * CW - 1) / (2 * CW), tiles_h = (d->Hi + 2 * CH - 1) / (2 * CH), tiles_d = (d->Di + 2 * CD - 1) / (2 * CD);
    const int64_t nblk = (int64_t)tiles_w * tiles_h * tiles_d * d->N;
    SP_REQUIRE(nblk < (1LL << 31), "tiled corrT: too many tiles");
    const int ciP = (d->Ci + 15) / 16 * 16;
    const size_t smem = smem_bytes<K>(MAXCO);
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(corrT_s2_tiled_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    dim3 grid((unsigned)nblk, (unsigned)(ciP / CIT));
    corrT_s2_tiled_kernel<K><<<grid, 32 * CD, smem_bytes<K>(d->Co), st>>>(*d, nPerG, ciP, tiles_w, tiles_h, tiles_d, src, wp, bias,
                                                                         scale, shift, dst);
    SP_LAUNCH_OK("corrT_s2_tiled_kernel");
    return 0;
}

static inline int sp_tiledT_launch(const SpConvDesc* d, int nPerG, const float* src, const float* wp, const float* bias,
                                   const float* scale, const float* shift, float* dst, cudaStream_t st) {
    if (d->k == 3) return sp_tiledT_launch_t<3>(d, nPerG, src, wp, bias, scale, shift, dst, st);
    return sp_tiledT_launch_t<2>(d, nPerG, src, wp, bias, scale, shift, dst, st);
}
