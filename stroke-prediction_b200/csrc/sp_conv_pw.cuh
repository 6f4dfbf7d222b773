// sp_conv_pw.cuh — pointwise (1x1x1) convolution tier: the HBM-bound channel mixes Cae3D.py:215,218 (decoder 16->16, 16->1 +
// sigmoid on the full 28x128x128 volume), Unet3D.py:50,52 (classify 16->32->2) and the step MLP Cae3D.py:126-132, forward,
// dgrad (the same kernel on the transposed weight pack) and wgrad.
//
// Forward / dgrad: thread = 2 voxels x DT destination channels; a warp reads 32 consecutive voxels = one contiguous run of
// the NDHWC tensor (128-bit loads), BatchNorm scale/shift applied on the fly, weights [Cs][DT] broadcast from shared
// memory.  Algorithmic bytes per voxel: 4 (Cs + Cd): the kernel is bound by HBM, not by its Cs*Cd FMAs.
// Wgrad: thread = (voxel lane, source-channel quad) with a 4 x 16 register tile, voxels walked grid-strided; lanes are
// summed by warp shuffles, warps through shared memory (fixed order), CTAs by wgrad_reduce_kernel: deterministic.
#pragma once
#include "sp_common.cuh"

namespace sp_pw {

constexpr int NT = 256;
constexpr int VPT = 2;                       // voxels per thread (forward)

// src: [rows][lds], dst: [rows][ldd].  w: packed [Cs][dP] (dP = Cd rounded up to 16).  Destination pass blockIdx.y covers
// channels [d0, d0 + DT).
template <int DT>
__global__ void __launch_bounds__(NT)
pw_fwd_kernel(const float* __restrict__ src, int lds, int Cs, float* __restrict__ dst, int ldd, int Cd, int64_t rows,
              int64_t rows_per_group, const float* __restrict__ w, int dP, const float* __restrict__ bias,
              const float* __restrict__ scale, const float* __restrict__ shift, int act, float alpha) {
    extern __shared__ __align__(16) float wsm[];      // [Cs][DT]
    const int d0 = blockIdx.y * DT;
    for (int i = threadIdx.x; i < Cs * DT; i += NT) {
        const int j = i % DT, c = i / DT;
        wsm[i] = (d0 + j < dP) ? w[(int64_t)c * dP + d0 + j] : 0.f;
    }
    __syncthreads();
    float b[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j) b[j] = (bias && d0 + j < Cd) ? bias[d0 + j] : 0.f;
    const bool vec_s = (Cs % 4 == 0) && (lds % 4 == 0);
    const bool vec_d = (DT % 4 == 0) && (ldd % 4 == 0) && (d0 + DT <= Cd);

    for (int64_t base = (int64_t)blockIdx.x * (NT * VPT); base < rows; base += (int64_t)gridDim.x * (NT * VPT)) {
        float acc[VPT][DT];
#pragma unroll
        for (int v = 0; v < VPT; ++v)
#pragma unroll
            for (int j = 0; j < DT; ++j) acc[v][j] = b[j];
        int64_t r[VPT];
        bool ok[VPT];
        const float* sc[VPT];
        const float* sh[VPT];
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            r[v] = base + v * NT + threadIdx.x;
            ok[v] = r[v] < rows;
            const int g = ok[v] ? (int)(r[v] / rows_per_group) : 0;
            sc[v] = scale ? scale + (int64_t)g * Cs : nullptr;
            sh[v] = scale ? shift + (int64_t)g * Cs : nullptr;
        }
        if (vec_s) {
            for (int c = 0; c < Cs; c += 4) {
                float4 x[VPT];
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    x[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok[v]) {
                        x[v] = sp_ldg_stream(reinterpret_cast<const float4*>(src + r[v] * lds + c));
                        if (scale) {
                            const float4 s4 = *reinterpret_cast<const float4*>(sc[v] + c), h4 = *reinterpret_cast<const float4*>(sh[v] + c);
                            x[v].x = fmaf(x[v].x, s4.x, h4.x); x[v].y = fmaf(x[v].y, s4.y, h4.y);
                            x[v].z = fmaf(x[v].z, s4.z, h4.z); x[v].w = fmaf(x[v].w, s4.w, h4.w);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float wv[DT];
                    if (DT % 4 == 0) {     // broadcast LDS.128
#pragma unroll
                        for (int j4 = 0; j4 < DT / 4; ++j4) {
                            const float4 t = reinterpret_cast<const float4*>(wsm + (c + u) * DT)[j4];
                            wv[j4 * 4 + 0] = t.x; wv[j4 * 4 + 1] = t.y; wv[j4 * 4 + 2] = t.z; wv[j4 * 4 + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < DT; ++j) wv[j] = wsm[(c + u) * DT + j];
                    }
#pragma unroll
                    for (int v = 0; v < VPT; ++v) {
                        const float xv = (u == 0) ? x[v].x : (u == 1) ? x[v].y : (u == 2) ? x[v].z : x[v].w;
#pragma unroll
                        for (int j = 0; j < DT; ++j) acc[v][j] = fmaf(xv, wv[j], acc[v][j]);
                    }
                }
            }
        } else {
            for (int c = 0; c < Cs; ++c) {
                float x[VPT];
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    x[v] = 0.f;
                    if (ok[v]) {
                        x[v] = src[r[v] * lds + c];
                        if (scale) x[v] = fmaf(x[v], sc[v][c], sh[v][c]);
                    }
                }
                const float* wr = wsm + c * DT;
#pragma unroll
                for (int j = 0; j < DT; ++j) {
                    const float wv = wr[j];
#pragma unroll
                    for (int v = 0; v < VPT; ++v) acc[v][j] = fmaf(x[v], wv, acc[v][j]);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            if (!ok[v]) continue;
            float* yp = dst + r[v] * ldd + d0;
            if (vec_d) {
#pragma unroll
                for (int j4 = 0; j4 < DT / 4; ++j4) {
                    float4 o;
                    o.x = sp_act_fwd(acc[v][j4 * 4 + 0], act, alpha); o.y = sp_act_fwd(acc[v][j4 * 4 + 1], act, alpha);
                    o.z = sp_act_fwd(acc[v][j4 * 4 + 2], act, alpha); o.w = sp_act_fwd(acc[v][j4 * 4 + 3], act, alpha);
                    reinterpret_cast<float4*>(yp)[j4] = o;
                }
            } else {
#pragma unroll
                for (int j = 0; j < DT; ++j)
                    if (d0 + j < Cd) yp[j] = sp_act_fwd(acc[v][j], act, alpha);
            }
        }
    }
}

// Cs = 16 source channels (the decoder's 16->16 / 16->1 mixes, Cae3D.py:215,218, and their dgrads): the four 128-bit loads
// of each of the thread's two voxels are issued before any arithmetic (8 loads in flight per thread; the runtime-Cs kernel
// above waits for every quad before loading the next).  A voxel's four quads share two 32-byte sectors: L1-allocating
// loads, so every sector crosses the L2 -> SM link once.
template <int DT>
__global__ void __launch_bounds__(NT)
pw16_fwd_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int Cd, int64_t rows,
                int64_t rows_per_group, const float* __restrict__ w, int dP, const float* __restrict__ bias,
                const float* __restrict__ scale, const float* __restrict__ shift, int act, float alpha) {
    constexpr int CS = 16;
    __shared__ __align__(16) float wsm[CS * DT];       // [Cs][DT]
    const int d0 = blockIdx.y * DT;
    for (int i = threadIdx.x; i < CS * DT; i += NT) {
        const int j = i % DT, c = i / DT;
        wsm[i] = (d0 + j < dP) ? w[(int64_t)c * dP + d0 + j] : 0.f;
    }
    __syncthreads();
    float b[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j) b[j] = (bias && d0 + j < Cd) ? bias[d0 + j] : 0.f;
    const bool vec_d = (DT % 4 == 0) && (ldd % 4 == 0) && (d0 + DT <= Cd);

    for (int64_t base = (int64_t)blockIdx.x * (NT * VPT); base < rows; base += (int64_t)gridDim.x * (NT * VPT)) {
        float4 x[VPT][CS / 4];
        int64_t r[VPT];
        bool ok[VPT];
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            r[v] = base + v * NT + threadIdx.x;
            ok[v] = r[v] < rows;
            const float4* p = reinterpret_cast<const float4*>(src + (ok[v] ? r[v] : 0) * lds);
#pragma unroll
            for (int q = 0; q < CS / 4; ++q) x[v][q] = ok[v] ? __ldg(p + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (scale) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
                const int g = ok[v] ? (int)(r[v] / rows_per_group) : 0;
                const float4* sc = reinterpret_cast<const float4*>(scale + (int64_t)g * CS);
                const float4* sh = reinterpret_cast<const float4*>(shift + (int64_t)g * CS);
#pragma unroll
                for (int q = 0; q < CS / 4; ++q) {
                    const float4 s4 = __ldg(sc + q), h4 = __ldg(sh + q);
                    x[v][q].x = fmaf(x[v][q].x, s4.x, h4.x); x[v][q].y = fmaf(x[v][q].y, s4.y, h4.y);
                    x[v][q].z = fmaf(x[v][q].z, s4.z, h4.z); x[v][q].w = fmaf(x[v][q].w, s4.w, h4.w);
                }
            }
        }
        float acc[VPT][DT];
#pragma unroll
        for (int v = 0; v < VPT; ++v)
#pragma unroll
            for (int j = 0; j < DT; ++j) acc[v][j] = b[j];
#pragma unroll
        for (int q = 0; q < CS / 4; ++q) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float wv[DT];
                if (DT % 4 == 0) {     // broadcast LDS.128
#pragma unroll
                    for (int j4 = 0; j4 < DT / 4; ++j4) {
                        const float4 t = reinterpret_cast<const float4*>(wsm + (q * 4 + u) * DT)[j4];
                        wv[j4 * 4 + 0] = t.x; wv[j4 * 4 + 1] = t.y; wv[j4 * 4 + 2] = t.z; wv[j4 * 4 + 3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < DT; ++j) wv[j] = wsm[(q * 4 + u) * DT + j];
                }
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    const float xv = (u == 0) ? x[v][q].x : (u == 1) ? x[v][q].y : (u == 2) ? x[v][q].z : x[v][q].w;
#pragma unroll
                    for (int j = 0; j < DT; ++j) acc[v][j] = fmaf(xv, wv[j], acc[v][j]);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            if (!ok[v]) continue;
            float* yp = dst + r[v] * ldd + d0;
            if (vec_d) {
#pragma unroll
                for (int j4 = 0; j4 < DT / 4; ++j4) {
                    float4 o;
                    o.x = sp_act_fwd(acc[v][j4 * 4 + 0], act, alpha); o.y = sp_act_fwd(acc[v][j4 * 4 + 1], act, alpha);
                    o.z = sp_act_fwd(acc[v][j4 * 4 + 2], act, alpha); o.w = sp_act_fwd(acc[v][j4 * 4 + 3], act, alpha);
                    reinterpret_cast<float4*>(yp)[j4] = o;
                }
            } else {
#pragma unroll
                for (int j = 0; j < DT; ++j)
                    if (d0 + j < Cd) yp[j] = sp_act_fwd(acc[v][j], act, alpha);
            }
        }
    }
}

// dW[co][ci] partial of this CTA: ws[blockIdx.x][co][ci].  a = I-side [rows][lda] (Ci channels), o = O-side [rows][ldo].
// NQP = threads per voxel (source-channel quads, power of two <= 16); output channels [co0, co0 + 16) per blockIdx.y.
template <int NQP>
__global__ void __launch_bounds__(NT)
pw_wgrad_kernel(const float* __restrict__ a, int lda, int Ci, const float* __restrict__ a_scale, const float* __restrict__ a_shift,
                const float* __restrict__ o, int ldo, int Co, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                int64_t rows, int64_t rows_per_group, float* __restrict__ ws) {
    constexpr int VL = NT / NQP;               // voxel lanes per CTA
    constexpr int WL = 32 / NQP;               // voxel lanes per warp
    __shared__ float red[NT / 32][NQP][64];
    const int q = threadIdx.x % NQP, vl = threadIdx.x / NQP;
    const int c0 = q * 4;
    const int co0 = blockIdx.y * 16;
    const bool vec_a = (Ci % 4 == 0) && (lda % 4 == 0);
    const bool vec_o = (ldo % 4 == 0) && (co0 + 16 <= Co);
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;

    auto load_row = [&](int64_t r, float* x, float* gz) {
        const int g = (int)(r / rows_per_group);
        if (vec_a && c0 < Ci) {
            const float4 t = sp_ldg_stream(reinterpret_cast<const float4*>(a + r * lda + c0));
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = (c0 + u < Ci) ? a[r * lda + c0 + u] : 0.f;
        }
        if (a_scale) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + u < Ci) x[u] = fmaf(x[u], a_scale[(int64_t)g * Ci + c0 + u], a_shift[(int64_t)g * Ci + c0 + u]);
        }
        if (vec_o) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const float4 t = *reinterpret_cast<const float4*>(o + r * ldo + co0 + j4 * 4);
                gz[j4 * 4 + 0] = t.x; gz[j4 * 4 + 1] = t.y; gz[j4 * 4 + 2] = t.z; gz[j4 * 4 + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) gz[j] = (co0 + j < Co) ? o[r * ldo + co0 + j] : 0.f;
        }
        if (o_scale) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (co0 + j < Co) gz[j] = fmaf(gz[j], o_scale[(int64_t)g * Co + co0 + j], o_shift[(int64_t)g * Co + co0 + j]);
        }
    };
    const int64_t step = (int64_t)gridDim.x * VL;
    int64_t r = (int64_t)blockIdx.x * VL + vl;
    for (; r + step < rows; r += 2 * step) {       // two rows in flight
        float x0[4], g0[16], x1[4], g1[16];
        load_row(r, x0, g0);
        load_row(r + step, x1, g1);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(x0[i], g0[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(x1[i], g1[j], acc[i][j]);
    }
    for (; r < rows; r += step) {
        float x0[4], g0[16];
        load_row(r, x0, g0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(x0[i], g0[j], acc[i][j]);
    }
    // voxel lanes of a warp (same quad: lanes q, q + NQP, ...) by xor-shuffles, then the warps through shared memory
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float v = acc[i][j];
#pragma unroll
            for (int off = NQP; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            acc[i][j] = v;
        }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < NQP) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) red[warp][lane][i * 16 + j] = acc[i][j];
    }
    __syncthreads();
    const int64_t wn = (int64_t)Co * Ci;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = threadIdx.x; i < NQP * 64; i += NT) {
        const int qq = i / 64, e = i % 64;
        const int ci = qq * 4 + e / 16, co = co0 + e % 16;
        if (ci < Ci && co < Co) {
            float v = red[0][qq][e];
#pragma unroll
            for (int wv = 1; wv < NT / 32; ++wv) v += red[wv][qq][e];
            wsp[(int64_t)co * Ci + ci] = v;
        }
    }
    (void)WL;
}

// Ci = 16 (four threads per voxel) with Co a multiple of 16 (COT = 16) or Co = 1 (COT = 1): the decoder's 16->16 / 16->1
// mixes (Cae3D.py:215,218) and Unet3D.py:50.  Every thread loads only ITS quad of the O-side row (the four threads of a
// voxel exchange them by shuffles) and keeps four rows in flight: twice the unique bytes in flight of pw_wgrad_kernel,
// which is latency-bound (ncu: 30 % of the DRAM roof, long-scoreboard stalls).
template <int COT>
__global__ void __launch_bounds__(NT, 2)
pw_wgrad16_kernel(const float* __restrict__ a, int lda, const float* __restrict__ a_scale, const float* __restrict__ a_shift,
                  const float* __restrict__ o, int ldo, int Co, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                  int64_t rows, int64_t rows_per_group, float* __restrict__ ws) {
    constexpr int NQP = 4, Ci = 16, VL = NT / NQP, R = 4, NE = 4 * COT;
    __shared__ float red[NT / 32][NQP][NE];
    const int q = threadIdx.x % NQP, vl = threadIdx.x / NQP;
    const int c0 = q * 4;
    const int co0 = blockIdx.y * 16;
    float acc[4][COT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < COT; ++j) acc[i][j] = 0.f;
    const int64_t step = (int64_t)gridDim.x * VL;
    // the trip count is uniform over the CTA (the quad exchange below is a full-warp shuffle)
    for (int64_t rbase = (int64_t)blockIdx.x * VL; rbase < rows; rbase += R * step) {
        const int64_t r = rbase + vl;
        float4 xa[R], og[R];
        bool ok[R];
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int64_t row = r + u * step;
            ok[u] = row < rows;
            xa[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            og[u] = xa[u];
            if (ok[u]) {
                xa[u] = sp_ldg_stream(reinterpret_cast<const float4*>(a + row * lda + c0));
                if (COT == 16) og[u] = sp_ldg_stream(reinterpret_cast<const float4*>(o + row * ldo + co0 + c0));
                else og[u].x = __ldg(o + row * ldo + co0);
            }
        }
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const int64_t row = r + u * step;
            float x[4] = {xa[u].x, xa[u].y, xa[u].z, xa[u].w};
            float gq[4] = {og[u].x, og[u].y, og[u].z, og[u].w};
            if (ok[u] && (a_scale || o_scale)) {
                const int g = (int)(row / rows_per_group);
                if (a_scale) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = fmaf(x[i], a_scale[(int64_t)g * Ci + c0 + i], a_shift[(int64_t)g * Ci + c0 + i]);
                }
                if (o_scale) {
                    if (COT == 16) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            gq[i] = fmaf(gq[i], o_scale[(int64_t)g * Co + co0 + c0 + i], o_shift[(int64_t)g * Co + co0 + c0 + i]);
                    } else {
                        gq[0] = fmaf(gq[0], o_scale[(int64_t)g * Co + co0], o_shift[(int64_t)g * Co + co0]);
                    }
                }
            }
            if (COT == 16) {
                float gz[16];
#pragma unroll
                for (int pq = 0; pq < 4; ++pq)
#pragma unroll
                    for (int i = 0; i < 4; ++i) gz[pq * 4 + i] = __shfl_sync(0xffffffffu, gq[i], pq, 4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < COT; ++j) acc[i][j] = fmaf(x[i], gz[j < 16 ? j : 0], acc[i][j]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][0] = fmaf(x[i], gq[0], acc[i][0]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < COT; ++j) {
            float v = acc[i][j];
#pragma unroll
            for (int off = NQP; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            acc[i][j] = v;
        }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < NQP) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < COT; ++j) red[warp][lane][i * COT + j] = acc[i][j];
    }
    __syncthreads();
    const int64_t wn = (int64_t)Co * Ci;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = threadIdx.x; i < NQP * NE; i += NT) {
        const int qq = i / NE, e = i % NE;
        const int ci = qq * 4 + e / COT, co = co0 + e % COT;
        if (co < Co) {
            float v = red[0][qq][e];
#pragma unroll
            for (int wv = 1; wv < NT / 32; ++wv) v += red[wv][qq][e];
            wsp[(int64_t)co * Ci + ci] = v;
        }
    }
}

static inline int nqp_for(int Ci) {
    int nq = (Ci + 3) / 4, p = 1;
    while (p < nq) p <<= 1;
    return p;
}
static inline int wgrad_grid_x(int64_t rows, int nqp) {
    const int vl = NT / nqp;
    int64_t gx = (int64_t)sp_num_sms() * 4;
    const int64_t need = sp_cdiv(rows, (int64_t)vl * 8);
    if (gx > need) gx = need;
    if (gx < 1) gx = 1;
    return (int)gx;
}

}  // namespace sp_pw

static inline bool sp_pw_disabled() {
    static int v = -1;   // SP_DISABLE_PW=1 forces the generic kernels
    if (v < 0) {
        const char* e = getenv("SP_DISABLE_PW");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// forward: Cs source channels -> Cd destination channels
static inline bool sp_pw_fwd_supported(const SpConvDesc* d, int Cs) {
    return d->k == 1 && d->s == 1 && d->pd == 0 && d->ph == 0 && d->pw == 0 && Cs <= 512 && !sp_pw_disabled();
}

static inline int sp_pw_fwd_launch(const float* src, int lds, int Cs, float* dst, int ldd, int Cd, int64_t rows, int64_t rows_per_group,
                                   const float* w, const float* bias, const float* scale, const float* shift, int act, float alpha,
                                   cudaStream_t st) {
    using namespace sp_pw;
    const int dP = (Cd + 15) / 16 * 16;
    int64_t gx = sp_cdiv(rows, NT * VPT);
    const int64_t cap = (int64_t)sp_num_sms() * 8;
    if (gx > cap) gx = cap;
    const bool fast16 = (Cs == 16) && (lds % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (fast16) {
        // same accumulation order per output as pw_fwd_kernel (c ascending, bias first): bit-identical results
        const int64_t cap16 = (int64_t)sp_num_sms() * 4;
        if (gx > cap16) gx = cap16;
        if (Cd > 4) {
            dim3 grid((unsigned)gx, (unsigned)(dP / 16));
            pw16_fwd_kernel<16><<<grid, NT, 0, st>>>(src, lds, dst, ldd, Cd, rows, rows_per_group, w, dP, bias, scale, shift, act, alpha);
        } else if (Cd > 1) {
            pw16_fwd_kernel<4><<<(unsigned)gx, NT, 0, st>>>(src, lds, dst, ldd, Cd, rows, rows_per_group, w, dP, bias, scale, shift, act, alpha);
        } else {
            pw16_fwd_kernel<1><<<(unsigned)gx, NT, 0, st>>>(src, lds, dst, ldd, Cd, rows, rows_per_group, w, dP, bias, scale, shift, act, alpha);
        }
        SP_LAUNCH_OK("pw16_fwd_kernel");
        return 0;
    }
    if (Cd > 4) {
        dim3 grid((unsigned)gx, (unsigned)(dP / 16));
        pw_fwd_kernel<16><<<grid, NT, (size_t)Cs * 16 * 4, st>>>(src, lds, Cs, dst, ldd, Cd, rows, rows_per_group, w, dP, bias, scale, shift, act, alpha);
    } else if (Cd > 1) {
        dim3 grid((unsigned)gx, 1);
        pw_fwd_kernel<4><<<grid, NT, (size_t)Cs * 4 * 4, st>>>(src, lds, Cs, dst, ldd, Cd, rows, rows_per_group, w, dP, bias, scale, shift, act, alpha);
    } else {
        dim3 grid((unsigned)gx, 1);
        pw_fwd_kernel<1><<<grid, NT, (size_t)Cs * 1 * 4, st>>>(src, lds, Cs, dst, ldd, Cd, rows, rows_per_group, w, dP, bias, scale, shift, act, alpha);
    }
    SP_LAUNCH_OK("pw_fwd_kernel");
    return 0;
}

static inline bool sp_pw_wgrad_supported(const SpConvDesc* d) {
    return d->k == 1 && d->s == 1 && d->pd == 0 && d->ph == 0 && d->pw == 0 && d->Ci <= 64 && !sp_pw_disabled();
}

static inline size_t sp_pw_wgrad_workspace_bytes(const SpConvDesc* d) {
    if (!sp_pw_wgrad_supported(d)) return 0;
    const int64_t rows = (int64_t)d->N * d->Do * d->Ho * d->Wo;
    return (size_t)sp_pw::wgrad_grid_x(rows, sp_pw::nqp_for(d->Ci)) * d->Co * d->Ci * sizeof(float);
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int chunks, int64_t wn, float* __restrict__ dw, float beta);

static inline int sp_pw_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                     const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                     cudaStream_t st) {
    using namespace sp_pw;
    const int64_t vox = (int64_t)d->Do * d->Ho * d->Wo;
    const int64_t rows = (int64_t)d->N * vox, rpg = (int64_t)nPerG * vox;
    const int nqp = nqp_for(d->Ci);
    const int gx = wgrad_grid_x(rows, nqp);
    dim3 grid(gx, (d->Co + 15) / 16);
    const bool al16 = ((reinterpret_cast<uintptr_t>(iside) | reinterpret_cast<uintptr_t>(oside)) & 15) == 0;
    if (d->Ci == 16 && d->ldi % 4 == 0 && al16 && ((d->Co % 16 == 0 && d->ldo % 4 == 0) || d->Co == 1)) {
        if (d->Co == 1)
            pw_wgrad16_kernel<1><<<grid, NT, 0, st>>>(iside, d->ldi, i_scale, i_shift, oside, d->ldo, d->Co, o_scale, o_shift, rows, rpg, ws);
        else
            pw_wgrad16_kernel<16><<<grid, NT, 0, st>>>(iside, d->ldi, i_scale, i_shift, oside, d->ldo, d->Co, o_scale, o_shift, rows, rpg, ws);
        SP_LAUNCH_OK("pw_wgrad16_kernel");
        const int64_t wn16 = (int64_t)d->Co * d->Ci;
        int64_t rb16 = (wn16 + 255) / 256;
        wgrad_reduce_kernel<<<(int)rb16, 256, 0, st>>>(ws, gx, wn16, dw, beta);
        SP_LAUNCH_OK("wgrad_reduce_kernel");
        return 0;
    }
#define SP_PWW(Q) pw_wgrad_kernel<Q><<<grid, NT, 0, st>>>(iside, d->ldi, d->Ci, i_scale, i_shift, oside, d->ldo, d->Co, o_scale, o_shift, rows, rpg, ws)
    switch (nqp) {
        case 1: SP_PWW(1); break;
        case 2: SP_PWW(2); break;
        case 4: SP_PWW(4); break;
        case 8: SP_PWW(8); break;
        default: SP_PWW(16); break;
    }
#undef SP_PWW
    SP_LAUNCH_OK("pw_wgrad_kernel");
    const int64_t wn = (int64_t)d->Co * d->Ci;
    int64_t rb = (wn + 255) / 256;
    if (rb > 148 * 16) rb = 148 * 16;
    wgrad_reduce_kernel<<<(int)rb, 256, 0, st>>>(ws, gx, wn, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}
