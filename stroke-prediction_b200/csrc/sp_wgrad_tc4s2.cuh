// sp_wgrad_tc4s2.cuh — tcgen05 / TMEM weight gradient of the STRIDE-2 3x3x3 layer with 9..16 input channels
// (Cae3D.py:48: Conv3d(16, 24, 3, stride 2, padding 1) — on the FFMA tier the largest single launch of the training step).
//
//   dW[co][ci][kd][kh][kw] = sum_{n,od,oh,ow} dZ[n,od,oh,ow,co] * X'[n, 2 od - 1 + kd, 2 oh - 1 + kh, 2 ow - 1 + kw, ci]
//
// Same formulation as sp_wgrad_tc4.cuh (voxels = contraction index, both operands MN-major, ONE M 128 x N 96 MMA per
// (16 voxels, kd), two round-to-nearest bf16 terms per operand); what the stride changes:
//   w  The 16 voxels of a K slab must be consecutive in shared memory, but consecutive outputs read every other input
//      column.  The input columns are therefore split by parity: a tile column holds the input columns u = 2 j + q of ONE
//      parity q, j = j0 .. j0 + 31, and the CTAs walk the tile columns of both parities.  Even columns (q = 0) carry the tap
//      kw = 1 (ow = j); odd columns (q = 1) carry kw = 2 (ow = j) and kw = 0 (ow = j + 1).  The three M row groups of
//      sp_wgrad_tc4.cuh become tap slots: slot 0 = kw 0 (dZ[j + 1], q = 1), slot 1 = kw 2 (dZ[j], q = 1), slot 2 = kw 1 (dZ[j],
//      q = 0); the slots of the other parity are stored as zeros, so ONE accumulator set serves both parities.
//   h  Output row r reads the input rows 2 r - 1 + kh: three CONSECUTIVE rows of the staged tile, starting at row 2 r — the
//      N groups (kh, term, ci half) keep their uniform stride.  A tile is 2 output rows = 5 input rows.
//   d  Output plane od reads the input planes 2 od - 1 + kd: the ring advances by TWO planes per step (three at the top of a
//      column); ring of 8 slots (see the staging schedule for the safety argument).
// More than 16 output channels (24 here) run as slices of 16 with their own launches, like the wider stride-1 layers.
// X' staging: thread = (column, channel half, one of the step's two new planes), all five rows of the plane loaded up front.
// Roles, barriers, drain and accuracy as in sp_wgrad_tc4.cuh.
#pragma once
#include "sp_wgrad_tc4.cuh"

namespace sp_wtc4s2 {

using namespace sp_tc;
using sp_tc2::mbar_arrive;
using sp_wtc::idesc_mn;
using sp_wtc4::mbar_wait_issuer;
using sp_wtc4::mbar_wait_warp;
using sp_wtc4::split8_rn2;
using sp_wtc4::tmem_ld32;

constexpr int TU = 32, THW = 2;                    // tile: 32 input columns of one parity x 2 output rows
constexpr int XH = 2 * THW + 1;                    // input rows of a tile
constexpr int ZW = TU + 2;                         // dZ columns a tile reads
constexpr int RS = TU * 16;                        // N-group stride (512)
constexpr int X_PLANE_B = XH * 4 * RS;             // [row][term][half][j] = 10240
constexpr int NBUF = 3;
constexpr int NSLOT = 8;
constexpr int X_REGION_B = NSLOT * X_PLANE_B;      // 81920
constexpr int PS = THW * TU * 16;                  // M-group stride (1024)
constexpr int A_BUF_B = 12 * PS;                   // [term][kw'][half][row][j] = 12288
constexpr int A_REGION_B = NBUF * A_BUF_B;
constexpr int NBLK = 3, BCOLS = 96;
constexpr int ACC_ROWS = 96, ACC_LD = 148;
constexpr int ACC_B = ACC_ROWS * ACC_LD * 4;
constexpr int MAXG = 16;
constexpr int COEF_B = MAXG * 32 * 4;
constexpr int W_EPI = 4, W_MMA = 3, W_STG_X = 4, W_STG_Z = 3, W_STG_G = W_STG_X + W_STG_Z, NGRP = NBUF;
constexpr int NTHREADS_S2 = (W_EPI + W_MMA + NGRP * W_STG_G) * 32;     // 896
constexpr int N_BARS = 2 * NBUF + 2 * NBLK;
constexpr size_t SMEM = (size_t)A_REGION_B + X_REGION_B + ACC_B + COEF_B + N_BARS * 8 + 16;
static_assert(SMEM <= 227 * 1024, "wgrad tc4 s2: shared memory");
static_assert(A_REGION_B + X_REGION_B >= (NBUF - 1) * A_BUF_B + 16 * PS, "the 16 M groups of the last buffer stay inside shared memory");
static_assert(W_STG_Z * 32 >= 2 * ZW, "staging roles");

struct ColGeo {
    int n, oh0, j0, q;
};
// The two parities of one tile column are NEIGHBOURING work items (q = item & 1): the CTAs that run side by side read the same
// 128-byte lines of X (two voxels of 64 B: one per parity) at the same time, so the second read hits L2.  Parity-major order
// re-read every line from DRAM: 1.98 GB per launch for 0.79 GB of operands (ncu, profiles/r02_ncu_tc_kernels_in_step.csv).
__device__ __forceinline__ ColGeo col_geo(int item, int tiles_w, int tiles_h) {
    ColGeo c;
    c.q = item & 1;
    int col = item >> 1;
    const int tw = col % tiles_w;
    col /= tiles_w;
    c.oh0 = (col % tiles_h) * THW;
    c.n = col / tiles_h;
    c.j0 = tw * TU;
    return c;
}

__global__ void __launch_bounds__(NTHREADS_S2, 1)
wgrad3_tc4s2_kernel(SpConvDesc d, int kk, int nPerG, int G, int tiles_w, int tiles_h, int total_cols, int drain_every, int isstride,
                    int osstride, const float* __restrict__ X, const float* __restrict__ i_scale, const float* __restrict__ i_shift,
                    const float* __restrict__ dZ, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                    float* __restrict__ ws, long long* __restrict__ prof, int nt) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* a_reg = smem_raw;                                              // [buf][term][kw'][half][row][j] x 16 B
    unsigned char* x_reg = smem_raw + A_REGION_B;                                 // [slot][row][term][half][j] x 16 B
    float* acc = reinterpret_cast<float*>(smem_raw + A_REGION_B + X_REGION_B);    // [(term_y, kw', co)][ACC_LD]
    float* coef = reinterpret_cast<float*>(smem_raw + A_REGION_B + X_REGION_B + ACC_B);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + A_REGION_B + X_REGION_B + ACC_B + COEF_B);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool pr = (prof != nullptr) && (blockIdx.x == 0);
    long long pw0 = 0, pw1 = 0, pwk = 0;

    for (int i = tid; i < ACC_ROWS * ACC_LD; i += NTHREADS_S2) acc[i] = 0.f;
    for (int i = tid; i < G * 32; i += NTHREADS_S2) {
        const int g = i >> 5, j = i & 31, c = j & 15;
        float v = (j < 16) ? 1.f : 0.f;
        if (i_scale && c < d.Ci) v = (j < 16) ? i_scale[(int64_t)g * isstride + c] : i_shift[(int64_t)g * isstride + c];
        coef[i] = v;
    }
    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(smem_u32(&bars[b]), W_STG_G);
            mbar_init(smem_u32(&bars[NBUF + b]), kk);                  // a_empty[b]: one commit per active issuer (kd < kk)
        }
        for (int b = 0; b < NBLK; ++b) {
            mbar_init(smem_u32(&bars[2 * NBUF + b]), 1);
            mbar_init(smem_u32(&bars[2 * NBUF + NBLK + b]), 3);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[NBUF]);
    const uint32_t t_full = smem_u32(&bars[2 * NBUF]), t_empty = smem_u32(&bars[2 * NBUF + NBLK]);

    const int ncols = (total_cols - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int Do = d.Do;
    const int nsteps = ncols * Do;
    const int ndrains = (nsteps + drain_every - 1) / drain_every;
    const int pad = kk - 2;                                             // k3: padding 1, k2: padding 0 (both: input index 2 o - pad + tap)
    const int PPC = 2 * Do + pad;                                       // planes per column in the ring's sequence numbering
    const int xh = 2 * THW + pad;                                       // input rows of a tile that are read

    if (warp >= W_EPI + W_MMA) {
        // =================================================================== staging: group grp takes the steps it = grp (mod NGRP)
        const int sw = warp - (W_EPI + W_MMA);
        const int grp = warp_uniform(sw / W_STG_G);
        const int wig = warp_uniform(sw % W_STG_G);
        const bool isx = wig < W_STG_X;
        const int st = (wig - (isx ? 0 : W_STG_X)) * 32 + lane;
        const bool vec_i = (d.ldi % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
        const bool vec_o = (d.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(dZ) & 15) == 0);
        const int64_t xs_n = (int64_t)d.Di * d.Hi * d.Wi * d.ldi, zs_n = (int64_t)Do * d.Ho * d.Wo * d.ldo;
        const int wx = st & 31, xhalf = (st >> 5) & 1, xp = st >> 6;       // X' role: plane xp of the step's pair, all five rows
        const bool zact = !isx && st < 2 * ZW;                             // dZ role: (channel half, dZ column), two rows
        const int zhalf = st >= ZW ? 1 : 0, zj = st - zhalf * ZW;
        for (int it = grp; it < nsteps; it += NGRP) {
            const int buf = it % NBUF, use = it / NBUF;
            const int cl = it / Do, od = it - cl * Do;
            const ColGeo cg = col_geo((int)blockIdx.x + cl * (int)gridDim.x, tiles_w, tiles_h);
            const int q = cg.q;
            bool waited = false;
            long long c1 = pr ? clock64() : 0;
            if (isx) {
                // Planes of this step in ring sequence numbers: column cl holds the numbers cl * PPC + 0 .. 2 Do (plane p of a
                // column is input depth p - 1), output plane od reads p = 2 od, 2 od + 1, 2 od + 2.  A step stages 2 od + 1 and
                // 2 od + 2, the top of a column also p = 0.  Ring of 8: the numbers a step writes replace those 8 below them; the
                // highest of these (window base + 2 - 8 = base - 6) was last read by the step three before (whose window reaches
                // its base + 2 = this base - 4 >= base - 6 ... base - 5 within a column, and across a column change — where the
                // base jumps by three — by the step three before as well): that step's MMAs have been waited for (a_empty).
                const int np = (od == 0 && kk == 3) ? 3 : 2;          // k2: windows do not overlap, planes 2 od and 2 od + 1
#pragma unroll 1
                for (int round = 0; round * 2 < np; ++round) {
                    const int pidx = round * 2 + xp;                    // newest plane first; index 2 only at od == 0 (plane 0)
                    const bool have = pidx < np;
                    const int pj = 2 * od + kk - 1 - pidx;
                    const int seq = cl * PPC + pj;
                    const int gd = pj - pad, gw = 2 * (cg.j0 + wx) + q, gh0 = 2 * cg.oh0 - pad;
                    const int ch = xhalf * 8;
                    const bool okp = have && gd >= 0 && gd < d.Di && gw < d.Wi;
                    const float* pp = X + (int64_t)cg.n * xs_n + (((int64_t)gd * d.Hi + gh0) * d.Wi + gw) * d.ldi + ch;
                    const int64_t rstep = (int64_t)d.Wi * d.ldi;
                    float4 ra[XH], rb[XH];
                    bool ok[XH];
#pragma unroll
                    for (int i = 0; i < XH; ++i) {
                        const int gh = gh0 + i;
                        ok[i] = okp && i < xh && gh >= 0 && gh < d.Hi;
                        ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        rb[i] = ra[i];
                        if (ok[i]) {
                            const float* p8 = pp + i * rstep;
                            if (vec_i && ch + 8 <= d.Ci) {
                                ra[i] = *reinterpret_cast<const float4*>(p8);
                                rb[i] = *reinterpret_cast<const float4*>(p8 + 4);
                            } else {
                                float e[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) e[j] = (ch + j < d.Ci) ? p8[j] : 0.f;
                                ra[i] = make_float4(e[0], e[1], e[2], e[3]);
                                rb[i] = make_float4(e[4], e[5], e[6], e[7]);
                            }
                        }
                    }
                    const float* cf = coef + (cg.n / nPerG) * 32 + ch;
                    const float4 s0 = *reinterpret_cast<const float4*>(cf), s1 = *reinterpret_cast<const float4*>(cf + 4);
                    const float4 h0 = *reinterpret_cast<const float4*>(cf + 16), h1 = *reinterpret_cast<const float4*>(cf + 20);
                    uint4 o1[XH], o2[XH];
#pragma unroll
                    for (int i = 0; i < XH; ++i) {
                        float v[8] = {ra[i].x, ra[i].y, ra[i].z, ra[i].w, rb[i].x, rb[i].y, rb[i].z, rb[i].w};
                        if (ok[i]) {
                            v[0] = fmaf(v[0], s0.x, h0.x); v[1] = fmaf(v[1], s0.y, h0.y); v[2] = fmaf(v[2], s0.z, h0.z); v[3] = fmaf(v[3], s0.w, h0.w);
                            v[4] = fmaf(v[4], s1.x, h1.x); v[5] = fmaf(v[5], s1.y, h1.y); v[6] = fmaf(v[6], s1.z, h1.z); v[7] = fmaf(v[7], s1.w, h1.w);
                        }
                        split8_rn2(v, o1[i], o2[i], nt);
                    }
                    if (!waited) {
                        if (pr) c1 = clock64();
                        mbar_wait_warp(a_empty + 8 * buf, (use & 1) ^ 1);       // the MMAs of step it - NBUF are done
                        waited = true;
                        if (pr) { const long long c2 = clock64(); pw0 += c2 - c1; c1 = c2; }
                    }
                    if (have) {
                        unsigned char* dp = x_reg + (seq % NSLOT) * X_PLANE_B + xhalf * RS + wx * 16;
#pragma unroll
                        for (int i = 0; i < XH; ++i) {
                            *reinterpret_cast<uint4*>(dp + i * 4 * RS) = o1[i];
                            *reinterpret_cast<uint4*>(dp + i * 4 * RS + 2 * RS) = o2[i];
                        }
                    }
                }
            } else {
                const float* zn = dZ + (int64_t)cg.n * zs_n;
                const int gz = cg.n / nPerG;
                const int gw = cg.j0 - 1 + zj, ch = zhalf * 8;           // A_kw'[j] = dZ[j + 1 - kw']
                const bool okc = zact && gw >= 0 && gw < d.Wo && ch < d.Co;
                const float* pp = zn + (((int64_t)od * d.Ho + cg.oh0) * d.Wo + gw) * d.ldo + ch;
                const int64_t rstep = (int64_t)d.Wo * d.ldo;
                float4 ra[THW], rb[THW];
                bool ok[THW];
#pragma unroll
                for (int i = 0; i < THW; ++i) {
                    ok[i] = okc && cg.oh0 + i < d.Ho;
                    ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    rb[i] = ra[i];
                    if (ok[i]) {
                        const float* p8 = pp + i * rstep;
                        if (vec_o && ch + 8 <= d.Co) {
                            ra[i] = *reinterpret_cast<const float4*>(p8);
                            rb[i] = *reinterpret_cast<const float4*>(p8 + 4);
                        } else {
                            float e[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) e[j] = (ch + j < d.Co) ? p8[j] : 0.f;
                            ra[i] = make_float4(e[0], e[1], e[2], e[3]);
                            rb[i] = make_float4(e[4], e[5], e[6], e[7]);
                        }
                    }
                }
                uint4 o1[THW], o2[THW];
#pragma unroll
                for (int i = 0; i < THW; ++i) {
                    float v[8] = {ra[i].x, ra[i].y, ra[i].z, ra[i].w, rb[i].x, rb[i].y, rb[i].z, rb[i].w};
                    if (ok[i] && o_scale) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (ch + j < d.Co) v[j] = fmaf(v[j], o_scale[(int64_t)gz * osstride + ch + j], o_shift[(int64_t)gz * osstride + ch + j]);
                    }
                    split8_rn2(v, o1[i], o2[i], nt);
                }
                if (pr) c1 = clock64();
                mbar_wait_warp(a_empty + 8 * buf, (use & 1) ^ 1);
                if (pr) { const long long c2 = clock64(); pw0 += c2 - c1; c1 = c2; }
                if (zact) {
                    // slot 0 = kw 0 <- dZ[j + 1] (k = zj - 2), slot 1 = kw 2 <- dZ[j] (k = zj - 1): odd columns; slot 2 = kw 1 <- dZ[j]
                    // (k = zj - 1): even columns; the other parity's slots get zeros (every (slot, k) is written every step)
                    unsigned char* dp = a_reg + buf * A_BUF_B + zhalf * PS;
                    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int slot = 0; slot < 3; ++slot) {
                        const int k = zj - (slot == 0 ? 2 : 1);
                        // k2: slot 2 = kw 0 <- dZ[j] (even columns), slot 1 = kw 1 <- dZ[j] (odd columns), slot 0 unused
                        const bool data = (slot == 2) ? (q == 0) : (q == 1 && (kk == 3 || slot == 1));
                        if (k >= 0 && k < TU) {
#pragma unroll
                            for (int i = 0; i < THW; ++i) {
                                unsigned char* p8 = dp + slot * 2 * PS + k * 16 + i * (TU * 16);
                                *reinterpret_cast<uint4*>(p8) = data ? o1[i] : z4;
                                *reinterpret_cast<uint4*>(p8 + 6 * PS) = data ? o2[i] : z4;
                            }
                        }
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * buf);
            if (pr) pwk += clock64() - c1;
        }
        if (pr && grp == 0 && wig == 0 && lane == 0) { prof[4] = pw0; prof[5] = pwk; }
        if (pr && grp == 0 && wig == W_STG_X && lane == 0) { prof[8] = pw0; prof[9] = pwk; }
    } else if (warp >= W_EPI) {
        // =================================================================== MMA issue: warp 4 + kd owns accumulator block kd
        {
            const int kd = warp_uniform(warp - W_EPI);
            const uint32_t a_base = smem_u32(a_reg), x_base = smem_u32(x_reg);
            const uint32_t dcol = tmem_base + (uint32_t)(kd * BCOLS);
            const uint32_t IDESC = idesc_mn(128, kk * 32);                  // N = (kh < kk, term, ci)
            bool fresh = true;
            int drains = 0;
            for (int it = 0; it < (kd < kk ? nsteps : 0); ++it) {
                const int buf = it % NBUF, use = it / NBUF;
                const int cl = it / Do, od = it - cl * Do;
                long long c0 = pr ? clock64() : 0;
                mbar_wait_issuer(a_full + 8 * buf, use & 1);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                if (fresh && drains > 0) mbar_wait_issuer(t_empty + 8 * kd, (drains - 1) & 1);
                long long c2 = pr ? clock64() : 0;
                pw1 += c2 - c1;
                tc_fence_after();
                const int slot = (cl * PPC + 2 * od + kd) % NSLOT;     // input plane 2 od - pad + kd of this column
                const uint64_t da0 = umma_desc(a_base + (uint32_t)(buf * A_BUF_B), 128, PS);
                const uint64_t db0 = umma_desc(x_base + (uint32_t)(slot * X_PLANE_B), 128, RS);
#pragma unroll
                for (int r = 0; r < THW; ++r) {
#pragma unroll
                    for (int ks = 0; ks < TU / 16; ++ks) {
                        const uint64_t da = da0 + (uint64_t)((r * TU * 16 + ks * 256) >> 4);
                        const uint64_t db = db0 + (uint64_t)((2 * r * 4 * RS + ks * 256) >> 4);      // input rows 2 r, 2 r + 1, 2 r + 2
                        umma_bf16_elect(dcol, da, db, IDESC, fresh ? 0u : 1u);
                        fresh = false;
                    }
                }
                umma_commit_elect(a_empty + 8 * buf);
                if ((it + 1) % drain_every == 0 || it == nsteps - 1) {
                    umma_commit_elect(t_full + 8 * kd);
                    fresh = true;
                    ++drains;
                }
                if (pr) pwk += clock64() - c2;
            }
            if (pr && kd == 0 && lane == 0) { prof[0] = pw0; prof[1] = pw1; prof[2] = pwk; prof[3] = nsteps; }
        }
    } else if (warp < 3) {
        // =================================================================== drain: warp w holds accumulator rows 32 w .. 32 w + 31
        float* arow = acc + (size_t)(warp * 32 + lane) * ACC_LD;
        for (int dr = 0; dr < ndrains; ++dr) {
#pragma unroll 1
            for (int kd = 0; kd < kk; ++kd) {
                long long c0 = pr ? clock64() : 0;
                mbar_wait_warp(t_full + 8 * kd, dr & 1);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                tc_fence_after();
#pragma unroll 1
                for (int kh = 0; kh < kk; ++kh) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kd * BCOLS + kh * 32), v);
                    if (kh == kk - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(t_empty + 8 * kd);
                    }
                    float4* ap = reinterpret_cast<float4*>(arow + kd * 48 + kh * 16);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        float4 a = ap[j4];
                        a.x += v[4 * j4] + v[16 + 4 * j4];
                        a.y += v[4 * j4 + 1] + v[17 + 4 * j4];
                        a.z += v[4 * j4 + 2] + v[18 + 4 * j4];
                        a.w += v[4 * j4 + 3] + v[19 + 4 * j4];
                        ap[j4] = a;
                    }
                }
                if (pr) pwk += clock64() - c1;
            }
        }
        if (pr && tid == 0) { prof[6] = pw0; prof[7] = pwk; }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
    // fold the two y terms (small first); this CTA's partial as [co][ci][kd][kh][slot] (slot -> kw in wgrad_reduce_s2_kernel)
    const int wn = d.Co * d.Ci * 27;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = tid; i < wn; i += NTHREADS_S2) {
        const int tap = i % 27, ci = (i / 27) % d.Ci, co = i / (27 * d.Ci);
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        const int col = kd * 48 + kh * 16 + ci;
        const int r0 = (kw * 2 + (co >> 3)) * 8 + (co & 7);
        wsp[i] = acc[(r0 + 48) * ACC_LD + col] + acc[r0 * ACC_LD + col];
    }
}

// dw[co0 + co][ci0 + ci][kd][kh][kw] (kk^3 taps) = beta * dw + sum_chunks ws[chunk][co][ci][(kd * 3 + kh) * 3 + slot];
// k3: slot 0 / 1 / 2 = kw 0 / 2 / 1;  k2: slot 2 / 1 = kw 0 / 1
__global__ void wgrad_reduce_s2_kernel(const float* __restrict__ ws, int chunks, int cos, int cs, int Ci, int co0, int ci0, int kk,
                                       float* __restrict__ dw, float beta) {
    const int k3 = kk * kk * kk;
    const int total = cos * cs * k3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int tap = i % k3, ci = (i / k3) % cs, co = i / (k3 * cs);
    const int kw = tap % kk, kh = (tap / kk) % kk, kd = tap / (kk * kk);
    const int slot = (kk == 3) ? (kw == 0 ? 0 : (kw == 2 ? 1 : 2)) : (kw == 0 ? 2 : 1);
    const int wn = cos * cs * 27;
    const int src = (co * cs + ci) * 27 + (kd * 3 + kh) * 3 + slot;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += ws[(int64_t)c * wn + src];
    float* o = dw + ((int64_t)(co0 + co) * Ci + ci0 + ci) * k3 + tap;
    *o = (beta != 0.f ? beta * *o : 0.f) + s;
}

struct Plan {
    int tiles_w, tiles_h, grid;
    int64_t total;      // work items: (tile column, parity), parity fastest
};
static inline Plan plan(const SpConvDesc* d) {
    Plan p;
    p.tiles_h = (d->Ho + THW - 1) / THW;
    p.tiles_w = ((d->Wi + 1) / 2 + TU - 1) / TU;                        // 32 input columns of one parity per tile (even columns: ceil(Wi / 2))
    p.total = 2 * (int64_t)p.tiles_w * p.tiles_h * d->N;
    p.grid = sp_num_sms();
    const int cap = sp_wtc4_grid_cap_ref();
    if (cap > 0 && p.grid > cap) p.grid = cap;
    if (p.grid > p.total) p.grid = (int)p.total;
    return p;
}

}  // namespace sp_wtc4s2

// Cae3D.py:48 (16 -> 24), Cae3D.py:59 (24 -> 32): 3x3x3 stride-2 padding-1 layers, and the decoder's ConvTranspose3d k2 s2 steps
// (Cae3D.py:193,204: 24 -> 24, 16 -> 16; kk = 2: two taps per axis, no overlap between windows), 9..32 channels, as slices of 16
static inline bool sp_tc4s2_wgrad_supported(const SpConvDesc* d, int G) {
    if (sp_tc4_wgrad_disabled() || G < 1 || G > sp_wtc4s2::MAXG) return false;
    if ((d->k != 3 && d->k != 2) || d->s != 2 || sp_tc_terms() == 0 || sp_tc_wgrad_disabled()) return false;
    const int pad = d->k - 2;                                            // Conv3d(k3, s2, p1) / ConvTranspose3d(k2, s2, p0)
    if (d->pd != pad || d->ph != pad || d->pw != pad) return false;
    // k2 layers with more than 16 channels stay on sp_conv_k2s2.cuh: measured 0.29 (k2s2 kernel) vs 0.37 ms (2 x 2 slices) on Cae3D.py:193
    if (d->k == 2 && (d->Ci > 16 || d->Co > 16)) return false;
    if (d->Ci <= 8 || d->Ci > 32 || d->Co <= 8 || d->Co > 32 || d->ldi % 4 != 0 || d->ldo % 4 != 0) return false;   // up to 2 x 2 slices
    if (d->Wi < 2 || d->Do < 4 || d->Wo < 16) return false;
    const sp_wtc4s2::Plan p = sp_wtc4s2::plan(d);
    return p.total >= 32 && p.total < (1LL << 31) / (2 * d->Do + 1);
}

static inline size_t sp_tc4s2_wgrad_workspace_bytes(const SpConvDesc* d) {
    return (size_t)sp_wtc4s2::plan(d).grid * 16 * 16 * 27 * sizeof(float);       // one output slice at a time (stream-ordered reuse)
}

static inline int sp_tc4s2_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                        const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                        cudaStream_t st, long long* prof = nullptr, int drain_every = 12) {   // 12 steps x 4 MMAs = the 48 accumulations per drain of sp_wgrad_tc4.cuh
    using namespace sp_wtc4s2;
    const Plan p = plan(d);
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(wgrad3_tc4s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        attr = true;
    }
    const int G = d->N / nPerG;
    const int nsi = (d->Ci + 15) / 16, nso = (d->Co + 15) / 16;
    for (int co = 0; co < nso; ++co)
        for (int c = 0; c < nsi; ++c) {
            SpConvDesc s = *d;
            s.Ci = (d->Ci - 16 * c < 16) ? d->Ci - 16 * c : 16;
            s.Co = (d->Co - 16 * co < 16) ? d->Co - 16 * co : 16;
            wgrad3_tc4s2_kernel<<<p.grid, NTHREADS_S2, SMEM, st>>>(s, d->k, nPerG, G, p.tiles_w, p.tiles_h, (int)p.total,
                                                                   drain_every, d->Ci, d->Co, iside + 16 * c, i_scale ? i_scale + 16 * c : nullptr,
                                                                   i_shift ? i_shift + 16 * c : nullptr, oside + 16 * co,
                                                                   o_scale ? o_scale + 16 * co : nullptr, o_shift ? o_shift + 16 * co : nullptr, ws,
                                                                   (co == 0 && c == 0) ? prof : nullptr, sp_tc_terms() == 1 ? 1 : 2);
            SP_LAUNCH_OK("wgrad3_tc4s2_kernel");
            const int wn = s.Co * s.Ci * d->k * d->k * d->k;
            wgrad_reduce_s2_kernel<<<(wn + 255) / 256, 256, 0, st>>>(ws, p.grid, s.Co, s.Ci, d->Ci, 16 * co, 16 * c, d->k, dw, beta);
            SP_LAUNCH_OK("wgrad_reduce_s2_kernel");
        }
    return 0;
}
