// sp_conv_k2s2.cuh — kernel 2, stride 2, padding 0 layers: the ConvTranspose3d up-sampling steps of the decoder
// (Cae3D.py:193,204: 24 -> 24 and 16 -> 16 channels), forward (sp_corrT), dgrad (sp_corr) and wgrad.
// Every I-side voxel belongs to exactly one 2x2x2 block = one O-side voxel: no halo, no overlap — the three ops are
// channel mixes with a tap-dependent weight, 256 MACs per I-side voxel against 64 + 8 bytes of traffic: HBM-bound (the
// halo-tile tiers spent 0.8 / 1.1 / 1.5 ms on the 16-channel layer for 0.15 ms of HBM time).
//   k2s2_up_kernel    sp_corrT: thread = (O voxel, I-side channel quad): 8 taps x 4 channels, the O voxel's channels are
//                     read once (broadcast over the quad lanes), weights [tap][co][ciP] broadcast from shared memory
//   k2s2_down_kernel  sp_corr : thread = (O voxel, O-side channel quad), the eight I voxels of the block streamed
//   k2s2_wgrad_kernel thread = (voxel lane, tap, I-side channel quad) with a 4 x 16 register tile: 8 taps x Ci/4 quads
//                     = every lane busy for 16 channels; lanes reduced in fixed order through shared memory
#pragma once
#include "sp_common.cuh"

#ifndef SP_K2S2_UP_MINB
// resident CTAs per SM asked of ptxas (ncu r02: latency-bound at 24 % occupancy with 91 registers; three CTAs: 0.697 -> 0.674 ms;
// the down kernel got slower with four: 0.588 -> 0.607 ms)
#define SP_K2S2_UP_MINB 3
#define SP_K2S2_DOWN_MINB 1
#define SP_K2S2_WGRAD_MINB 2
#endif
namespace sp_k2s2 {

constexpr int NT = 256;

// src = O-side [N][Do][Ho][Wo][ldo] (Co channels, optional affine), dst = I-side [N][2Do][2Ho][2Wo][ldi] (Ci channels).
// wt: Wt[tap][co][ciP].  Shared memory: 8 * Co * ciP floats.
template <int SQ>      // source channel quads (Co / 4): all loads of a voxel are issued before the first FMA
__global__ void __launch_bounds__(NT, SP_K2S2_UP_MINB)
k2s2_up_kernel(SpConvDesc d, int nPerG, int ciP, const float* __restrict__ src, const float* __restrict__ wt,
               const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ dst) {
    extern __shared__ __align__(16) float wsm[];
    for (int i = threadIdx.x; i < 8 * d.Co * ciP; i += NT) wsm[i] = wt[i];
    __syncthreads();
    const int CQ = d.Ci / 4;
    // 32-bit index arithmetic (the launcher checks items < 2^31): 64-bit div / mod would cost as much as the FMAs
    const unsigned items = (unsigned)d.N * (unsigned)(d.Do * d.Ho * d.Wo) * (unsigned)CQ;
    for (unsigned item = blockIdx.x * NT + threadIdx.x; item < items; item += gridDim.x * NT) {
        const int q = (int)(item % (unsigned)CQ);
        const unsigned vu = item / (unsigned)CQ;
        const int64_t v = vu;
        const int ow = (int)(vu % (unsigned)d.Wo); unsigned r = vu / (unsigned)d.Wo;
        const int oh = (int)(r % (unsigned)d.Ho); r /= (unsigned)d.Ho;
        const int od = (int)(r % (unsigned)d.Do);
        const int n = (int)(r / (unsigned)d.Do);
        const int g = n / nPerG;
        const float* sp = src + v * d.ldo;
        float4 acc[8];
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b4 = *reinterpret_cast<const float4*>(bias + q * 4);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = b4;
        float4 sq[SQ];
#pragma unroll
        for (int k = 0; k < SQ; ++k) sq[k] = __ldg(reinterpret_cast<const float4*>(sp + k * 4));
        if (scale) {
#pragma unroll
            for (int k = 0; k < SQ; ++k) {
                const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + (int64_t)g * d.Co + k * 4));
                const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + (int64_t)g * d.Co + k * 4));
                sq[k].x = fmaf(sq[k].x, sc.x, sh.x); sq[k].y = fmaf(sq[k].y, sc.y, sh.y);
                sq[k].z = fmaf(sq[k].z, sc.z, sh.z); sq[k].w = fmaf(sq[k].w, sc.w, sh.w);
            }
        }
#pragma unroll
        for (int k = 0; k < SQ; ++k) {
            const float sv[4] = {sq[k].x, sq[k].y, sq[k].z, sq[k].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float4 w = *reinterpret_cast<const float4*>(wsm + ((t * d.Co + k * 4 + u) * ciP + q * 4));
                    acc[t].x = fmaf(sv[u], w.x, acc[t].x); acc[t].y = fmaf(sv[u], w.y, acc[t].y);
                    acc[t].z = fmaf(sv[u], w.z, acc[t].z); acc[t].w = fmaf(sv[u], w.w, acc[t].w);
                }
            }
        }
        float* dn = dst + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi + q * 4;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int id = 2 * od + (t >> 2), ih = 2 * oh + ((t >> 1) & 1), iw = 2 * ow + (t & 1);
            float4 o = acc[t];
            o.x = sp_act_fwd(o.x, d.act, d.alpha); o.y = sp_act_fwd(o.y, d.act, d.alpha);
            o.z = sp_act_fwd(o.z, d.act, d.alpha); o.w = sp_act_fwd(o.w, d.act, d.alpha);
            *reinterpret_cast<float4*>(dn + (((int64_t)id * d.Hi + ih) * d.Wi + iw) * d.ldi) = o;
        }
    }
}

// src = I-side (Ci channels, optional affine), dst = O-side (Co channels).  wc: Wc[tap][ci][coP].  Shared memory: 8 * Ci * coP floats.
template <int SQ>      // source channel quads (Ci / 4)
__global__ void __launch_bounds__(NT, SP_K2S2_DOWN_MINB)
k2s2_down_kernel(SpConvDesc d, int nPerG, int coP, const float* __restrict__ src, const float* __restrict__ wc,
                 const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ dst) {
    extern __shared__ __align__(16) float wsm[];
    for (int i = threadIdx.x; i < 8 * d.Ci * coP; i += NT) wsm[i] = wc[i];
    __syncthreads();
    const int CQ = d.Co / 4;
    // 32-bit index arithmetic (the launcher checks items < 2^31): 64-bit div / mod would cost as much as the FMAs
    const unsigned items = (unsigned)d.N * (unsigned)(d.Do * d.Ho * d.Wo) * (unsigned)CQ;
    for (unsigned item = blockIdx.x * NT + threadIdx.x; item < items; item += gridDim.x * NT) {
        const int q = (int)(item % (unsigned)CQ);
        const unsigned vu = item / (unsigned)CQ;
        const int64_t v = vu;
        const int ow = (int)(vu % (unsigned)d.Wo); unsigned r = vu / (unsigned)d.Wo;
        const int oh = (int)(r % (unsigned)d.Ho); r /= (unsigned)d.Ho;
        const int od = (int)(r % (unsigned)d.Do);
        const int n = (int)(r / (unsigned)d.Do);
        const int g = n / nPerG;
        const float* sn = src + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;
        float4 acc[2];                                   // two chains (even / odd taps), summed at the end
        acc[0] = bias ? *reinterpret_cast<const float4*>(bias + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int id = 2 * od + (t >> 2), ih = 2 * oh + ((t >> 1) & 1), iw = 2 * ow + (t & 1);
            const float* sp = sn + (((int64_t)id * d.Hi + ih) * d.Wi + iw) * d.ldi;
            float4 sq[SQ];
#pragma unroll
            for (int k = 0; k < SQ; ++k) sq[k] = __ldg(reinterpret_cast<const float4*>(sp + k * 4));
            if (scale) {
#pragma unroll
                for (int k = 0; k < SQ; ++k) {
                    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + (int64_t)g * d.Ci + k * 4));
                    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + (int64_t)g * d.Ci + k * 4));
                    sq[k].x = fmaf(sq[k].x, sc.x, sh.x); sq[k].y = fmaf(sq[k].y, sc.y, sh.y);
                    sq[k].z = fmaf(sq[k].z, sc.z, sh.z); sq[k].w = fmaf(sq[k].w, sc.w, sh.w);
                }
            }
#pragma unroll
            for (int k = 0; k < SQ; ++k) {
                const float sv[4] = {sq[k].x, sq[k].y, sq[k].z, sq[k].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 w = *reinterpret_cast<const float4*>(wsm + ((t * d.Ci + k * 4 + u) * coP + q * 4));
                    float4& a = acc[t & 1];
                    a.x = fmaf(sv[u], w.x, a.x); a.y = fmaf(sv[u], w.y, a.y); a.z = fmaf(sv[u], w.z, a.z); a.w = fmaf(sv[u], w.w, a.w);
                }
            }
        }
        float4 o;
        o.x = sp_act_fwd(acc[0].x + acc[1].x, d.act, d.alpha); o.y = sp_act_fwd(acc[0].y + acc[1].y, d.act, d.alpha);
        o.z = sp_act_fwd(acc[0].z + acc[1].z, d.act, d.alpha); o.w = sp_act_fwd(acc[0].w + acc[1].w, d.act, d.alpha);
        *reinterpret_cast<float4*>(dst + v * d.ldo + q * 4) = o;
    }
}

// ws[cta][co][ci][8].  blockIdx.y = pass over 16 output channels.  LPV = 8 * Ci/4 threads per voxel.
__global__ void __launch_bounds__(NT, SP_K2S2_WGRAD_MINB)
k2s2_wgrad_kernel(SpConvDesc d, int nPerG, const float* __restrict__ iside, const float* __restrict__ i_scale,
                  const float* __restrict__ i_shift, const float* __restrict__ oside, const float* __restrict__ o_scale,
                  const float* __restrict__ o_shift, float* __restrict__ ws) {
    __shared__ float red[48][65];
    const int CIQ = d.Ci / 4, LPV = 8 * CIQ, VL = NT / LPV;
    const int item = threadIdx.x % LPV, vl = threadIdx.x / LPV;
    const bool active = vl < VL;
    const int t = item / CIQ, q = item % CIQ;
    const int co0 = blockIdx.y * 16;
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
    // a voxel lane walks whole O-side rows (n, od, oh): two divisions per row, pointer increments inside
    const int nrows = d.N * d.Do * d.Ho;
    const int tw = t & 1;
    if (active) {
        for (int row = blockIdx.x * VL + vl; row < nrows; row += gridDim.x * VL) {
            const int oh = row % d.Ho;
            const int rr = row / d.Ho;
            const int od = rr % d.Do, n = rr / d.Do;
            const int g = n / nPerG;
            const int id = 2 * od + (t >> 2), ih = 2 * oh + ((t >> 1) & 1);
            const float* xp = iside + ((((int64_t)n * d.Di + id) * d.Hi + ih) * d.Wi + tw) * d.ldi + q * 4;
            const float* op = oside + (int64_t)row * d.Wo * d.ldo + co0;
            float4 isc = make_float4(1.f, 1.f, 1.f, 1.f), ish = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i_scale) {
                isc = __ldg(reinterpret_cast<const float4*>(i_scale + (int64_t)g * d.Ci + q * 4));
                ish = __ldg(reinterpret_cast<const float4*>(i_shift + (int64_t)g * d.Ci + q * 4));
            }
            auto load = [&](int ow, float4& x, float* gz) {
                x = sp_ldg_stream(reinterpret_cast<const float4*>(xp + (int64_t)(2 * ow) * d.ldi));
                if (i_scale) {
                    x.x = fmaf(x.x, isc.x, ish.x); x.y = fmaf(x.y, isc.y, ish.y); x.z = fmaf(x.z, isc.z, ish.z); x.w = fmaf(x.w, isc.w, ish.w);
                }
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (co0 + j4 * 4 < d.Co) {
                        o = __ldg(reinterpret_cast<const float4*>(op + (int64_t)ow * d.ldo + j4 * 4));
                        if (o_scale) {
                            const float4 sc = __ldg(reinterpret_cast<const float4*>(o_scale + (int64_t)g * d.Co + co0 + j4 * 4));
                            const float4 sh = __ldg(reinterpret_cast<const float4*>(o_shift + (int64_t)g * d.Co + co0 + j4 * 4));
                            o.x = fmaf(o.x, sc.x, sh.x); o.y = fmaf(o.y, sc.y, sh.y); o.z = fmaf(o.z, sc.z, sh.z); o.w = fmaf(o.w, sc.w, sh.w);
                        }
                    }
                    gz[j4 * 4] = o.x; gz[j4 * 4 + 1] = o.y; gz[j4 * 4 + 2] = o.z; gz[j4 * 4 + 3] = o.w;
                }
            };
            int ow = 0;
            for (; ow + 1 < d.Wo; ow += 2) {               // two voxels in flight
                float4 x0, x1;
                float g0[16], g1[16];
                load(ow, x0, g0);
                load(ow + 1, x1, g1);
                const float xa[4] = {x0.x, x0.y, x0.z, x0.w}, xb[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(xa[i], g0[j], acc[i][j]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(xb[i], g1[j], acc[i][j]);
            }
            if (ow < d.Wo) {
                float4 x0;
                float g0[16];
                load(ow, x0, g0);
                const float xa[4] = {x0.x, x0.y, x0.z, x0.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(xa[i], g0[j], acc[i][j]);
            }
        }
    }
    // voxel lanes folded in fixed order through shared memory
    for (int w = 0; w < VL; ++w) {
        if (active && vl == w) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 16; ++j) red[item][i * 16 + j] = (w == 0 ? 0.f : red[item][i * 16 + j]) + acc[i][j];
        }
        __syncthreads();
    }
    const int64_t wn = (int64_t)d.Co * d.Ci * 8;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = threadIdx.x; i < LPV * 64; i += NT) {
        const int it = i / 64, e = i % 64;
        const int tt = it / CIQ, qq = it % CIQ;
        const int ci = qq * 4 + e / 16, co = co0 + e % 16;
        if (co < d.Co) wsp[((int64_t)co * d.Ci + ci) * 8 + tt] = red[it][e];
    }
}

}  // namespace sp_k2s2

static inline bool sp_k2s2_disabled() {
    static int v = -1;   // SP_DISABLE_K2S2=1 sends these layers back to the halo-tile tiers
    if (v < 0) {
        const char* e = getenv("SP_DISABLE_K2S2");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static inline bool sp_k2s2_supported(const SpConvDesc* d) {
    if (d->k != 2 || d->s != 2 || d->pd || d->ph || d->pw || sp_k2s2_disabled()) return false;
    if (d->Di != 2 * d->Do || d->Hi != 2 * d->Ho || d->Wi != 2 * d->Wo) return false;
    if ((d->Ci != 16 && d->Ci != 24) || (d->Co != 16 && d->Co != 24) || d->ldi % 4 || d->ldo % 4) return false;
    const int64_t ov = (int64_t)d->N * d->Do * d->Ho * d->Wo;
    return ov >= 4096 && ov * 8 < (1LL << 31);
}
static inline bool sp_k2s2_aligned(const void* a, const void* b) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}
static inline int sp_k2s2_grid(int64_t items) {
    int64_t gx = (items + sp_k2s2::NT - 1) / sp_k2s2::NT;
    const int64_t cap = (int64_t)sp_num_sms() * 8;
    if (gx > cap) gx = cap;
    return (int)(gx < 1 ? 1 : gx);
}

static inline int sp_k2s2_up_launch(const SpConvDesc* d, int nPerG, const float* src, const float* wt, const float* bias,
                                    const float* scale, const float* shift, float* dst, cudaStream_t st) {
    using namespace sp_k2s2;
    const int ciP = (d->Ci + 15) / 16 * 16;
    const size_t smem = (size_t)8 * d->Co * ciP * 4;
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(k2s2_up_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 24 * 32 * 4));
        SP_CUDA(cudaFuncSetAttribute(k2s2_up_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 24 * 32 * 4));
        attr = true;
    }
    const int64_t items = (int64_t)d->N * d->Do * d->Ho * d->Wo * (d->Ci / 4);
    if (d->Co == 16) k2s2_up_kernel<4><<<sp_k2s2_grid(items), NT, smem, st>>>(*d, nPerG, ciP, src, wt, bias, scale, shift, dst);
    else k2s2_up_kernel<6><<<sp_k2s2_grid(items), NT, smem, st>>>(*d, nPerG, ciP, src, wt, bias, scale, shift, dst);
    SP_LAUNCH_OK("k2s2_up_kernel");
    return 0;
}

static inline int sp_k2s2_down_launch(const SpConvDesc* d, int nPerG, const float* src, const float* wc, const float* bias,
                                      const float* scale, const float* shift, float* dst, cudaStream_t st) {
    using namespace sp_k2s2;
    const int coP = (d->Co + 15) / 16 * 16;
    const size_t smem = (size_t)8 * d->Ci * coP * 4;
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(k2s2_down_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 24 * 32 * 4));
        SP_CUDA(cudaFuncSetAttribute(k2s2_down_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 24 * 32 * 4));
        attr = true;
    }
    const int64_t items = (int64_t)d->N * d->Do * d->Ho * d->Wo * (d->Co / 4);
    if (d->Ci == 16) k2s2_down_kernel<4><<<sp_k2s2_grid(items), NT, smem, st>>>(*d, nPerG, coP, src, wc, bias, scale, shift, dst);
    else k2s2_down_kernel<6><<<sp_k2s2_grid(items), NT, smem, st>>>(*d, nPerG, coP, src, wc, bias, scale, shift, dst);
    SP_LAUNCH_OK("k2s2_down_kernel");
    return 0;
}

static inline int sp_k2s2_wgrad_grid(const SpConvDesc* d) {
    const int vl = sp_k2s2::NT / (8 * (d->Ci / 4));
    int64_t gx = ((int64_t)d->N * d->Do * d->Ho + vl - 1) / vl;
    const int64_t cap = (int64_t)sp_num_sms() * 2;
    if (gx > cap) gx = cap;
    return (int)(gx < 1 ? 1 : gx);
}
static inline size_t sp_k2s2_wgrad_workspace_bytes(const SpConvDesc* d) {
    if (!sp_k2s2_supported(d)) return 0;
    return (size_t)sp_k2s2_wgrad_grid(d) * d->Co * d->Ci * 8 * sizeof(float);
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int chunks, int64_t wn, float* __restrict__ dw, float beta);

static inline int sp_k2s2_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                       const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                       cudaStream_t st) {
    using namespace sp_k2s2;
    const int gx = sp_k2s2_wgrad_grid(d);
    dim3 grid(gx, (d->Co + 15) / 16);
    k2s2_wgrad_kernel<<<grid, NT, 0, st>>>(*d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, ws);
    SP_LAUNCH_OK("k2s2_wgrad_kernel");
    const int64_t wn = (int64_t)d->Co * d->Ci * 8;
    wgrad_reduce_kernel<<<(int)((wn + 255) / 256), 256, 0, st>>>(ws, gx, wn, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}
