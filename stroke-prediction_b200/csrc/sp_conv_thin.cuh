// sp_conv_thin.cuh — 3x3x3 stride-1 layers whose I-side has 1..3 channels: the first convolution of every network
// (Cae3D.py:41 on the 1-channel masks, Enc3DCtp's 3-channel input Cae3D.py:153-158, Unet3D.py:19 block1 on CBV/TTD).
// They move one 16-channel activation tensor through HBM for 27*Ci*16 MACs per voxel (AI 7..20 F/B: HBM / LSU-bound), but
// the channel-quad tiers pad Ci to 4 and spend 4x the FMAs and scalar loads on them (0.94 / 0.75 / 1.34 ms forward / dgrad /
// wgrad of the CAE's first layer against 0.11 ms of HBM time).  Three specialised kernels:
//   thin_fwd_kernel<CI>   forward: thread = 4 consecutive w voxels x 16 output channels, scalar input planes in shared memory
//   thin_bwd_kernel       dgrad (flipped correlation 16 -> Ci channels): thread = 4 voxels x 1 channel, 16-channel halo tile in
//                         shared memory, weights of one (kd, kh) row in registers.  (Off the training path when the layer is the first
//                         unit of a network and its input needs no gradient: sp_bn_grads_from_wgrad)
//   thin_wgrad_kernel<CI> wgrad: thread = (row of output voxels, output-channel quad, ci) with all 27 taps in registers (108
//                         accumulators), sliding 3x3x3 input window, O-side quad prefetched one voxel ahead
#pragma once
#include "sp_common.cuh"

#ifndef SP_THIN_FWD_MINB
#define SP_THIN_FWD_MINB 2
#define SP_THIN_FWD1_MINB 2   // three CTAs per SM (80 registers, 176 B of spills): 0.687 -> 0.739 ms, not taken
#define SP_THIN_BWD_MINB 2
#define SP_THIN_WGRAD_MINB 1
#endif
namespace sp_thin {

constexpr int TW = 32, TH = 8, TD = 4;           // forward output tile
constexpr int XW = TW + 2, XH = TH + 2, XD = TD + 2;
constexpr int XRS = 36;                          // padded row (floats): 16-byte aligned quads at w0 = 4 * wq
constexpr int NT = 256;

// ---------------------------------------------------------------------------------------------------------------- forward
// wp: packed [tap][ci][coP]; flip != 0 reads tap 26 - t.  blockIdx.y = pass over 16 output channels.
template <int CI>
__global__ void __launch_bounds__(NT, (CI == 1) ? SP_THIN_FWD1_MINB : SP_THIN_FWD_MINB)
thin_fwd_kernel(SpConvDesc d, int nPerG, int coP, int tiles_w, int tiles_h, int tiles_d, const float* __restrict__ src,
                const float* __restrict__ wp, int flip, const float* __restrict__ bias, const float* __restrict__ scale,
                const float* __restrict__ shift, float* __restrict__ dst) {
    __shared__ __align__(16) float xs[CI * XD * XH * XRS];
    __shared__ __align__(16) float wsm[27 * CI * 16];
    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th_ = t % tiles_h; t /= tiles_h;
    const int td_ = t % tiles_d;
    const int n = t / tiles_d;
    const int ow0 = tw * TW, oh0 = th_ * TH, od0 = td_ * TD;
    const int co0 = blockIdx.y * 16;
    const int g = n / nPerG;
    const int id0 = od0 - d.pd, ih0 = oh0 - d.ph, iw0 = ow0 - d.pw;
    const float* srcn = src + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;

    for (int i = threadIdx.x; i < XD * XH * XW * CI; i += NT) {
        const int c = i % CI;
        int r = i / CI;
        const int wx = r % XW; r /= XW;
        const int hy = r % XH;
        const int dz = r / XH;
        const int gd = id0 + dz, gh = ih0 + hy, gw = iw0 + wx;
        float v = 0.f;
        if (gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi) {
            v = __ldg(srcn + (((int64_t)gd * d.Hi + gh) * d.Wi + gw) * d.ldi + c);
            if (scale) v = fmaf(v, scale[(int64_t)g * d.Ci + c], shift[(int64_t)g * d.Ci + c]);
        }
        xs[((c * XD + dz) * XH + hy) * XRS + wx] = v;
    }
    for (int i = threadIdx.x; i < 27 * CI * 16; i += NT) {
        const int j = i % 16, c = (i / 16) % CI, tap = i / (16 * CI);
        const int ts = flip ? 26 - tap : tap;
        wsm[i] = (co0 + j < coP) ? wp[((int64_t)ts * d.Ci + c) * coP + co0 + j] : 0.f;
    }
    __syncthreads();

    const int wq = threadIdx.x & 7, lth = (threadIdx.x >> 3) & 7, ltd = threadIdx.x >> 6;
    float2 acc[4][8];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[v][j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < CI; ++c) {
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const float* row = xs + ((c * XD + ltd + kd) * XH + lth + kh) * XRS + wq * 4;
                const float4 xa = *reinterpret_cast<const float4*>(row);
                const float2 xb = *reinterpret_cast<const float2*>(row + 4);
                const float x[6] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y};
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float4* wr = reinterpret_cast<const float4*>(wsm + (((kd * 3 + kh) * 3 + kw) * CI + c) * 16);
                    const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
                    const float2 wv[8] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w),
                                          make_float2(w2.x, w2.y), make_float2(w2.z, w2.w), make_float2(w3.x, w3.y), make_float2(w3.z, w3.w)};
#pragma unroll
                    for (int v = 0; v < 4; ++v)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[v][j] = sp_ffma2(x[v + kw], wv[j], acc[v][j]);
                }
            }
        }
    }
    const int od = od0 + ltd, oh = oh0 + lth;
    if (od >= d.Do || oh >= d.Ho) return;
    float b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) b[j] = (bias && co0 + j < d.Co) ? bias[co0 + j] : 0.f;
    const bool vec = (d.ldo % 4 == 0) && (co0 + 16 <= d.Co);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int ow = ow0 + wq * 4 + v;
        if (ow >= d.Wo) continue;
        float* yp = dst + ((((int64_t)n * d.Do + od) * d.Ho + oh) * d.Wo + ow) * d.ldo + co0;
        float o[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[2 * j] = sp_act_fwd(acc[v][j].x + b[2 * j], d.act, d.alpha);
            o[2 * j + 1] = sp_act_fwd(acc[v][j].y + b[2 * j + 1], d.act, d.alpha);
        }
        if (vec) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) reinterpret_cast<float4*>(yp)[j4] = make_float4(o[4 * j4], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (co0 + j < d.Co) yp[j] = o[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------ dgrad
// f: the FLIPPED correlation geometry (f.Ci = source channels <= 16, f.Co = 1..3 destination channels, pads k-1-p).
// wt: Wt[tap][co][ciP] of the original layer read with the flipped tap index = Wc[tap'][ci_f][co_f].
constexpr int BW = 32, BH = 8, BD = 2;           // dgrad output tile: 128 threads x 4 voxels
constexpr int BXW = BW + 2, BXH = BH + 2, BXD = BD + 2;
constexpr int BNT = 128;
constexpr int BRW = BXW | 1;                     // odd row stride (float4 units): the 8 rows of a quarter-warp hit 8 bank groups
constexpr int BPLANE = BXD * BXH * BRW + 1;      // quad plane stride, 16 bytes off a multiple of 128
constexpr size_t BWD_SMEM = (size_t)4 * BPLANE * 16;

__global__ void __launch_bounds__(BNT, SP_THIN_BWD_MINB)
thin_bwd_kernel(SpConvDesc f, int nPerG, int ciP, int tiles_w, int tiles_h, int tiles_d, const float* __restrict__ src,
                const float* __restrict__ wt, const float* __restrict__ bias, const float* __restrict__ scale,
                const float* __restrict__ shift, float* __restrict__ dst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* xs = reinterpret_cast<float4*>(smem_raw);                // [quad][dz][hy][BRW]
    __shared__ __align__(16) float wsm[3 * 27 * 16];                 // [co_f][tap'][ci_f]
    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th_ = t % tiles_h; t /= tiles_h;
    const int td_ = t % tiles_d;
    const int n = t / tiles_d;
    const int ow0 = tw * BW, oh0 = th_ * BH, od0 = td_ * BD;
    const int g = n / nPerG;
    const int id0 = od0 - f.pd, ih0 = oh0 - f.ph, iw0 = ow0 - f.pw;
    const float* srcn = src + (int64_t)n * f.Di * f.Hi * f.Wi * f.ldi;
    const bool vec = (f.ldi % 4 == 0);

    constexpr int NITEM = BXD * BXH * BXW * 4;
    if (!scale && vec && f.Ci == 16) {
        // plain gradient tile (Conv3d dgrad): asynchronous 16-byte copies, zero fill outside the volume — all of a thread's
        // copies are in flight at once
        for (int i = threadIdx.x; i < NITEM; i += BNT) {
            const int q = i & 3;
            int r = i >> 2;
            const int wx = r % BXW; r /= BXW;
            const int hy = r % BXH;
            const int dz = r / BXH;
            const int gd = id0 + dz, gh = ih0 + hy, gw = iw0 + wx;
            const bool in = gd >= 0 && gd < f.Di && gh >= 0 && gh < f.Hi && gw >= 0 && gw < f.Wi;
            const float* p = in ? srcn + (((int64_t)gd * f.Hi + gh) * f.Wi + gw) * f.ldi + q * 4 : srcn;
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(xs + q * BPLANE + (dz * BXH + hy) * BRW + wx);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(sa), "l"(p), "r"(in ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        constexpr int UB = 4;
        for (int i0 = threadIdx.x; i0 < NITEM; i0 += BNT * UB) {
            float e[UB][4];
            int so[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int i = i0 + u * BNT;
                so[u] = -1;
#pragma unroll
                for (int k = 0; k < 4; ++k) e[u][k] = 0.f;
                if (i >= NITEM) continue;
                const int q = i & 3;
                int r = i >> 2;
                const int wx = r % BXW; r /= BXW;
                const int hy = r % BXH;
                const int dz = r / BXH;
                so[u] = q * BPLANE + (dz * BXH + hy) * BRW + wx;
                const int gd = id0 + dz, gh = ih0 + hy, gw = iw0 + wx;
                if (gd >= 0 && gd < f.Di && gh >= 0 && gh < f.Hi && gw >= 0 && gw < f.Wi && q * 4 < f.Ci) {
                    const float* p = srcn + (((int64_t)gd * f.Hi + gh) * f.Wi + gw) * f.ldi + q * 4;
                    if (vec && q * 4 + 4 <= f.Ci) {
                        const float4 v = sp_ldg_stream(reinterpret_cast<const float4*>(p));
                        e[u][0] = v.x; e[u][1] = v.y; e[u][2] = v.z; e[u][3] = v.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) e[u][k] = (q * 4 + k < f.Ci) ? p[k] : 0.f;
                    }
                    if (scale) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (q * 4 + k < f.Ci) e[u][k] = fmaf(e[u][k], scale[(int64_t)g * f.Ci + q * 4 + k], shift[(int64_t)g * f.Ci + q * 4 + k]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UB; ++u)
                if (so[u] >= 0) xs[so[u]] = make_float4(e[u][0], e[u][1], e[u][2], e[u][3]);
        }
    }
    for (int i = threadIdx.x; i < f.Co * 27 * 16; i += BNT) {
        const int c = i % 16, tap = (i / 16) % 27, co = i / (16 * 27);
        wsm[i] = (c < f.Ci) ? wt[((int64_t)(26 - tap) * f.Ci + c) * ciP + co] : 0.f;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    // lane -> row (lane & 7), column group (lane >> 3); warp -> depth plane, column-group half
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lth = lane & 7, wq = (warp & 1) * 4 + (lane >> 3), ltd = warp >> 1;
    const int od = od0 + ltd, oh = oh0 + lth;
    for (int co = 0; co < f.Co; ++co) {
        float2 acc[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[v] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll 1
            for (int kh = 0; kh < 3; ++kh) {
                float4 w[3][4];                                   // weights of this (kd, kh) row: [kw][quad]
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        w[kw][q] = *reinterpret_cast<const float4*>(wsm + ((co * 27 + (kd * 3 + kh) * 3 + kw) * 16 + q * 4));
                const float4* row = xs + ((ltd + kd) * BXH + lth + kh) * BRW + wq * 4;
#pragma unroll
                for (int p = 0; p < 6; ++p) {
                    const float4 x0 = row[p], x1 = row[BPLANE + p], x2 = row[2 * BPLANE + p], x3 = row[3 * BPLANE + p];
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int v = p - kw;
                        if (v < 0 || v > 3) continue;
                        float2 a = acc[v];
                        a.x = fmaf(x0.x, w[kw][0].x, a.x); a.y = fmaf(x0.y, w[kw][0].y, a.y);
                        a.x = fmaf(x0.z, w[kw][0].z, a.x); a.y = fmaf(x0.w, w[kw][0].w, a.y);
                        a.x = fmaf(x1.x, w[kw][1].x, a.x); a.y = fmaf(x1.y, w[kw][1].y, a.y);
                        a.x = fmaf(x1.z, w[kw][1].z, a.x); a.y = fmaf(x1.w, w[kw][1].w, a.y);
                        a.x = fmaf(x2.x, w[kw][2].x, a.x); a.y = fmaf(x2.y, w[kw][2].y, a.y);
                        a.x = fmaf(x2.z, w[kw][2].z, a.x); a.y = fmaf(x2.w, w[kw][2].w, a.y);
                        a.x = fmaf(x3.x, w[kw][3].x, a.x); a.y = fmaf(x3.y, w[kw][3].y, a.y);
                        a.x = fmaf(x3.z, w[kw][3].z, a.x); a.y = fmaf(x3.w, w[kw][3].w, a.y);
                        acc[v] = a;
                    }
                }
            }
        }
        if (od < f.Do && oh < f.Ho) {
            const float bv = bias ? bias[co] : 0.f;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int ow = ow0 + wq * 4 + v;
                if (ow < f.Wo)
                    dst[((((int64_t)n * f.Do + od) * f.Ho + oh) * f.Wo + ow) * f.ldo + co] = sp_act_fwd((acc[v].x + acc[v].y) + bv, f.act, f.alpha);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------ wgrad
// thread = (output row (n, od, oh), output-channel quad q, input channel ci): LPV = 4 * CIP lanes per row (CIP = Ci
// rounded up to a power of two), 32 / LPV rows per warp (consecutive oh: their input rows overlap in L1).
// ws[cta][co][ci][27].
template <int CIP>
__global__ void __launch_bounds__(NT, SP_THIN_WGRAD_MINB)
thin_wgrad_kernel(SpConvDesc d, int nPerG, int64_t total_rows, const float* __restrict__ X, const float* __restrict__ i_scale,
                  const float* __restrict__ i_shift, const float* __restrict__ dZ, const float* __restrict__ o_scale,
                  const float* __restrict__ o_shift, float* __restrict__ ws) {
    constexpr int LPV = 4 * CIP, RPW = 32 / LPV, RPC = (NT / 32) * RPW;       // rows per warp / per CTA pass
    __shared__ float red[LPV][108];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane & 3, ci = (lane >> 2) % CIP, rw = lane / LPV;
    const bool act_lane = (ci < d.Ci) && (q * 4 < d.Co);
    const bool vec_o = (d.ldo % 4 == 0) && (q * 4 + 4 <= d.Co);
    float2 acc[27][2];
#pragma unroll
    for (int tp = 0; tp < 27; ++tp)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[tp][j] = make_float2(0.f, 0.f);

    for (int64_t row0 = (int64_t)blockIdx.x * RPC; row0 < total_rows; row0 += (int64_t)gridDim.x * RPC) {
        const int64_t row = row0 + warp * RPW + rw;
        if (row >= total_rows || !act_lane) continue;
        int64_t r = row;
        const int oh = (int)(r % d.Ho); r /= d.Ho;
        const int od = (int)(r % d.Do);
        const int n = (int)(r / d.Do);
        const int g = n / nPerG;
        const float sc = i_scale ? i_scale[(int64_t)g * d.Ci + ci] : 1.f, sh = i_scale ? i_shift[(int64_t)g * d.Ci + ci] : 0.f;
        float osc[4] = {1.f, 1.f, 1.f, 1.f}, osh[4] = {0.f, 0.f, 0.f, 0.f};
        if (o_scale) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (q * 4 + j < d.Co) { osc[j] = o_scale[(int64_t)g * d.Co + q * 4 + j]; osh[j] = o_shift[(int64_t)g * d.Co + q * 4 + j]; }
        }
        // the nine input rows (kd, kh) of this output row; invalid rows (zero padding) point nowhere
        const float* xr[9];
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int gd = od - d.pd + kd, gh = oh - d.ph + kh;
                xr[kd * 3 + kh] = (gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi)
                                      ? X + ((((int64_t)n * d.Di + gd) * d.Hi + gh) * d.Wi) * d.ldi + ci
                                      : nullptr;
            }
        auto ldx = [&](int rr, int gw) -> float {
            const float* p = xr[rr];
            if (p == nullptr || gw < 0 || gw >= d.Wi) return 0.f;
            return fmaf(__ldg(p + (int64_t)gw * d.ldi), sc, sh);
        };
        const float* zrow = dZ + ((((int64_t)n * d.Do + od) * d.Ho + oh) * d.Wo) * d.ldo + q * 4;
        auto ldz = [&](int ow, float* z) {
            if (vec_o) {
                const float4 v = sp_ldg_stream(reinterpret_cast<const float4*>(zrow + (int64_t)ow * d.ldo));
                z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) z[j] = (q * 4 + j < d.Co) ? zrow[(int64_t)ow * d.ldo + j] : 0.f;
            }
        };
        // O-side quads are fetched a group of four voxels ahead (HBM / L2 latency), the nine new input values one voxel
        // ahead (L1 latency); the FMAs of a voxel never wait for a load issued in the same iteration.
        auto ldz4 = [&](int ow0, float (*z)[4]) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (ow0 + u < d.Wo) ldz(ow0 + u, z[u]);
                else { z[u][0] = 0.f; z[u][1] = 0.f; z[u][2] = 0.f; z[u][3] = 0.f; }
            }
        };
        float win[9][3], xn[9];                                   // x'[row][ow - pw + kw]; xn = the value entering next
#pragma unroll
        for (int rr = 0; rr < 9; ++rr) {
            win[rr][0] = 0.f;
            win[rr][1] = ldx(rr, -d.pw);
            win[rr][2] = ldx(rr, 1 - d.pw);
            xn[rr] = ldx(rr, 2 - d.pw);
        }
        auto proc4 = [&](int ow0, float (*zq)[4]) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ow = ow0 + u;
                if (ow >= d.Wo) break;
                float2 z2[2];
                z2[0] = make_float2(o_scale ? fmaf(zq[u][0], osc[0], osh[0]) : zq[u][0], o_scale ? fmaf(zq[u][1], osc[1], osh[1]) : zq[u][1]);
                z2[1] = make_float2(o_scale ? fmaf(zq[u][2], osc[2], osh[2]) : zq[u][2], o_scale ? fmaf(zq[u][3], osc[3], osh[3]) : zq[u][3]);
#pragma unroll
                for (int rr = 0; rr < 9; ++rr) {
                    win[rr][0] = win[rr][1];
                    win[rr][1] = win[rr][2];
                    win[rr][2] = xn[rr];
                    xn[rr] = ldx(rr, ow + 3 - d.pw);
                }
#pragma unroll
                for (int rr = 0; rr < 9; ++rr)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                        for (int jp = 0; jp < 2; ++jp) acc[rr * 3 + kw][jp] = sp_ffma2(win[rr][kw], z2[jp], acc[rr * 3 + kw][jp]);
            }
        };
        float za[4][4], zb[4][4];
        ldz4(0, za);
        for (int ow0 = 0; ow0 < d.Wo; ow0 += 8) {
            ldz4(ow0 + 4, zb);
            proc4(ow0, za);
            ldz4(ow0 + 8, za);
            proc4(ow0 + 4, zb);
        }
    }
    // rows of a warp (lanes l, l + LPV, ...) by xor-shuffles, then the warps through shared memory in fixed order
#pragma unroll
    for (int tp = 0; tp < 27; ++tp)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float2 v = acc[tp][j];
#pragma unroll
            for (int off = LPV; off < 32; off <<= 1) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
                v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
            }
            acc[tp][j] = v;
        }
    for (int wv = 0; wv < NT / 32; ++wv) {
        if (warp == wv && lane < LPV) {
#pragma unroll
            for (int tp = 0; tp < 27; ++tp)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    red[lane][tp * 4 + j] = (wv == 0 ? 0.f : red[lane][tp * 4 + j]) + ((j & 1) ? acc[tp][j >> 1].y : acc[tp][j >> 1].x);
        }
        __syncthreads();
    }
    const int wn = d.Co * d.Ci * 27;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = threadIdx.x; i < LPV * 108; i += NT) {
        const int l = i / 108, e = i % 108;
        const int qq = l & 3, cc = (l >> 2) % CIP;
        const int tp = e / 4, co = qq * 4 + e % 4;
        if (cc < d.Ci && co < d.Co) wsp[((int64_t)co * d.Ci + cc) * 27 + tp] = red[l][e];
    }
}

}  // namespace sp_thin

static inline bool sp_thin_disabled() {
    static int v = -1;   // SP_DISABLE_THIN=1 sends the thin layers back to the channel-quad tiers
    if (v < 0) {
        const char* e = getenv("SP_DISABLE_THIN");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// forward: d = correlation geometry with Ci <= 3
static inline bool sp_thin_fwd_supported(const SpConvDesc* d) {
    if (d->k != 3 || d->s != 1 || d->Ci > 3 || sp_thin_disabled()) return false;
    if (d->pd < 0 || d->ph < 0 || d->pw < 0) return false;
    return (int64_t)d->Do * d->Ho * d->Wo >= 2048 && d->Wo >= 8 && d->Ho >= 4;
}
static inline int sp_thin_fwd_launch(const SpConvDesc* d, int nPerG, const float* src, const float* wp, int flip, const float* bias,
                                     const float* scale, const float* shift, float* dst, cudaStream_t st) {
    using namespace sp_thin;
    const int coP = (d->Co + 15) / 16 * 16;
    const int tiles_w = (d->Wo + TW - 1) / TW, tiles_h = (d->Ho + TH - 1) / TH, tiles_d = (d->Do + TD - 1) / TD;
    const int64_t tiles = (int64_t)tiles_w * tiles_h * tiles_d * d->N;
    SP_REQUIRE(tiles < (1LL << 31), "thin fwd: too many tiles");
    dim3 grid((unsigned)tiles, (unsigned)(coP / 16));
    if (d->Ci == 1) thin_fwd_kernel<1><<<grid, NT, 0, st>>>(*d, nPerG, coP, tiles_w, tiles_h, tiles_d, src, wp, flip, bias, scale, shift, dst);
    else if (d->Ci == 2) thin_fwd_kernel<2><<<grid, NT, 0, st>>>(*d, nPerG, coP, tiles_w, tiles_h, tiles_d, src, wp, flip, bias, scale, shift, dst);
    else thin_fwd_kernel<3><<<grid, NT, 0, st>>>(*d, nPerG, coP, tiles_w, tiles_h, tiles_d, src, wp, flip, bias, scale, shift, dst);
    SP_LAUNCH_OK("thin_fwd_kernel");
    return 0;
}

// dgrad: f = flipped correlation geometry (f.Ci <= 16 source channels, f.Co <= 3 destination channels)
static inline bool sp_thin_bwd_supported(const SpConvDesc* f) {
    if (f->k != 3 || f->s != 1 || f->Co > 3 || f->Ci > 16 || f->Ci < 4 || sp_thin_disabled()) return false;
    if (f->pd < 0 || f->ph < 0 || f->pw < 0) return false;
    return (int64_t)f->Do * f->Ho * f->Wo >= 2048 && f->Wo >= 8 && f->Ho >= 4;
}
static inline int sp_thin_bwd_launch(const SpConvDesc* f, int nPerG, const float* src, const float* wt, const float* bias,
                                     const float* scale, const float* shift, float* dst, cudaStream_t st) {
    using namespace sp_thin;
    const int ciP = (f->Co + 15) / 16 * 16;      // padded I-side channel count of the original layer
    const int tiles_w = (f->Wo + BW - 1) / BW, tiles_h = (f->Ho + BH - 1) / BH, tiles_d = (f->Do + BD - 1) / BD;
    const int64_t tiles = (int64_t)tiles_w * tiles_h * tiles_d * f->N;
    SP_REQUIRE(tiles < (1LL << 31), "thin bwd: too many tiles");
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(thin_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
        attr = true;
    }
    thin_bwd_kernel<<<(unsigned)tiles, BNT, BWD_SMEM, st>>>(*f, nPerG, ciP, tiles_w, tiles_h, tiles_d, src, wt, bias, scale, shift, dst);
    SP_LAUNCH_OK("thin_bwd_kernel");
    return 0;
}

static inline bool sp_thin_wgrad_supported(const SpConvDesc* d) {
    if (d->k != 3 || d->s != 1 || d->Ci > 3 || d->Co > 16 || sp_thin_disabled()) return false;
    return (int64_t)d->Do * d->Ho * d->Wo >= 2048 && d->Wo >= 8;
}
static inline int sp_thin_wgrad_grid(const SpConvDesc* d) {
    const int cip = d->Ci == 3 ? 4 : d->Ci;
    const int rpc = (sp_thin::NT / 32) * (32 / (4 * cip));
    const int64_t rows = (int64_t)d->N * d->Do * d->Ho;
    int64_t gx = sp_cdiv(rows, rpc);
    if (gx > sp_num_sms()) gx = sp_num_sms();
    return (int)gx;
}
static inline size_t sp_thin_wgrad_workspace_bytes(const SpConvDesc* d) {
    if (!sp_thin_wgrad_supported(d)) return 0;
    return (size_t)sp_thin_wgrad_grid(d) * d->Co * d->Ci * 27 * sizeof(float);
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int chunks, int64_t wn, float* __restrict__ dw, float beta);

static inline int sp_thin_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                       const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                       cudaStream_t st) {
    using namespace sp_thin;
    const int gx = sp_thin_wgrad_grid(d);
    const int64_t rows = (int64_t)d->N * d->Do * d->Ho;
    if (d->Ci == 1) thin_wgrad_kernel<1><<<gx, NT, 0, st>>>(*d, nPerG, rows, iside, i_scale, i_shift, oside, o_scale, o_shift, ws);
    else if (d->Ci == 2) thin_wgrad_kernel<2><<<gx, NT, 0, st>>>(*d, nPerG, rows, iside, i_scale, i_shift, oside, o_scale, o_shift, ws);
    else thin_wgrad_kernel<4><<<gx, NT, 0, st>>>(*d, nPerG, rows, iside, i_scale, i_shift, oside, o_scale, o_shift, ws);
    SP_LAUNCH_OK("thin_wgrad_kernel");
    const int64_t wn = (int64_t)d->Co * d->Ci * 27;
    wgrad_reduce_kernel<<<(int)((wn + 255) / 256), 256, 0, st>>>(ws, gx, wn, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}
