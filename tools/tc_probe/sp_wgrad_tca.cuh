// sp_wgrad_tca.cuh — EXPERIMENT (probe only, not part of libstroke_b200.so): the tcgen05 weight gradient of sp_wgrad_tc.cuh /
// sp_wgrad_tc24.cuh with the dZ operand in TENSOR MEMORY.  Correct (same rel-L2 as the shipped kernels), but on the B200 it is
// SLOWER: 4.59 ms vs 3.25 ms (16 channels), 1.11 vs 0.90 ms (24 channels) — the per-K-step producer -> issuer -> commit
// handshake (tcgen05.st + wait::st + mbarrier round trips, ~750 cycles per K step and producer warp) costs more than the
// shared-memory A fetch it removes.  Kept for the next round (profiles/r01_wgrad_tca_probe.log).
//
// The first generation is bound by shared-memory bandwidth, not by the tensor pipe (ncu: pipe active 50 %): every one of the
// 27 MMAs of a K step (16 voxels) re-reads the same 2-4 KB dZ tile (A operand) next to its own 1.5-2.25 KB of X' (B operand).
// tcgen05.mma takes A from TMEM (".ts" form, K-major, two bf16 per 32-bit column, lane = row), so here
//   * the staging warps copy the raw fp32 dZ tile [voxel][channel] to shared memory (no split);
//   * three PRODUCER warps (TMEM lane quarters 0..2 = bf16 terms 1..3; lane = output channel) read their channel's 16 voxels
//     of a K step, form their term of the exact three-way split, pack voxel pairs and tcgen05.st them into a ring of K-step
//     slots (8 TMEM columns each) behind the accumulators;
//   * the issuer warps multiply A[tmem slot] x B[smem descriptor] (M = 128: rows 32 t + co), and release the slot by a commit.
// Everything else is the first generation: MN-major X' with (kh, ci) as N, (kd, kw) accumulator blocks, a ring of input
// planes walked along the depth axis, periodic drains into fp32 shared-memory sums with the three y terms folded smallest
// first.  NG = 2: 9..16 channels, nine blocks x 48 columns in one pass; NG = 3: 17..24 channels, blocks 0..4 / 5..8 x 72
// columns in two passes.
#pragma once
#include "../../stroke-prediction_b200/csrc/sp_wgrad_tc24.cuh"

namespace sp_wtca {

using namespace sp_tc;
using sp_tc2::mbar_arrive;
using sp_tc2::split8_trunc3;

constexpr int TWW = 32, THW = 4, XW = TWW + 2, XH = THW + 2;
constexpr int RS = XW * 16;
constexpr int KSTEPS = TWW * THW / 16;             // 8 K steps per tile
// Warp layout (TMEM lane quarter = warp % 4): 0-2 drain, 4-6 and 8-10 two producer sets (even / odd K steps), issuers in warps
// 3, 7, 11, 12, ..., then the staging warps.
constexpr int W_FIXED = 12;

template <int NG>
struct Cfg {
    static constexpr int CP = 8 * NG;                          // padded channels per side
    static constexpr int X_ROW_B = NG * RS;
    static constexpr int X_PLANE_B = XH * X_ROW_B;
    static constexpr int NSLOT = (NG == 2) ? 6 : 4;            // input-plane ring (4: a column start waits for the previous step)
    static constexpr int X_TERM_B = NSLOT * X_PLANE_B;
    static constexpr int X_REGION_B = 3 * X_TERM_B;            // 117504 for both
    static constexpr int Z_BUF_B = TWW * THW * CP * 4;         // raw fp32 dZ tile [voxel][CP]
    static constexpr int BCOLS = 3 * CP;                       // N of one MMA = (kh, ci)
    static constexpr int MAXBLK = (NG == 2) ? 9 : 5;
    static constexpr int DCOLS = MAXBLK * BCOLS;               // 432 / 360 accumulator columns
    static constexpr int ASLOTS = (512 - DCOLS) / 8;           // K-step slots of the A ring: 10 / 19
    static constexpr int ACC_LD = DCOLS + 4;
    static constexpr int ACC_B = CP * ACC_LD * 4;
    static constexpr int SCR_LD = BCOLS + 4;
    static constexpr int SCR_B = CP * SCR_LD * 4;
    static constexpr int W_STG = (NG == 2) ? 6 : 8;            // staging warps (register budget: 768 / 704 threads)
    static constexpr int W_STG0 = W_FIXED + MAXBLK - 3;        // first staging warp
    static constexpr int NTHREADS = (W_STG0 + W_STG) * 32;
    static constexpr int N_BARS = 2 + 2 + 2 * MAXBLK + 2 * ASLOTS;
    static constexpr size_t SMEM = (size_t)X_REGION_B + 2 * Z_BUF_B + ACC_B + SCR_B + N_BARS * 8 + 16;
    static constexpr int XP_ITEMS = XH * XW * NG;              // 8-channel items of one input plane
    static constexpr int NZ_ITEMS = TWW * THW * NG;            // 8-channel items of the dZ tile
};
static_assert(Cfg<2>::SMEM <= 227 * 1024 && Cfg<3>::SMEM <= 227 * 1024, "wgrad tca: shared memory");

constexpr int PER_R = 4;

// D = f32, A = bf16 K-major (TMEM), B = bf16 MN-major (bit 16), M = 128
__host__ __device__ constexpr uint32_t idesc_ts(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void drain_bar() { asm volatile("bar.sync 1, 96;" ::: "memory"); }

// BCOLS consecutive TMEM columns in 8-column pieces, one wait
template <int NCOL>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
    uint32_t r[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; c += 8)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[c]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]), "=r"(r[c + 6]), "=r"(r[c + 7])
                     : "r"(taddr + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < NCOL; ++i) v[i] = __uint_as_float(r[i]);
}

// blocks [b0, b0 + nblk) of the nine (kd, kw) accumulator blocks (b = kd * 3 + kw)
template <int NG>
__global__ void __launch_bounds__(Cfg<NG>::NTHREADS, 1)
wgrad3_tca_kernel(SpConvDesc d, int nPerG, int tiles_w, int tiles_h, int total_cols, int drain_every, int b0, int nblk,
                  const float* __restrict__ X, const float* __restrict__ i_scale, const float* __restrict__ i_shift,
                  const float* __restrict__ dZ, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                  float* __restrict__ ws, long long* __restrict__ prof) {
    using C = Cfg<NG>;
    constexpr int CP = C::CP, BCOLS = C::BCOLS, MAXBLK = C::MAXBLK, ASLOTS = C::ASLOTS, NSLOT = C::NSLOT;
    constexpr int W_STG0 = C::W_STG0, NSTG = C::W_STG * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* x_reg = smem_raw;                                          // [term][slot][row][group][w] x 16 B
    float* z_reg = reinterpret_cast<float*>(smem_raw + C::X_REGION_B);        // [2][voxel][CP] raw fp32
    float* acc = reinterpret_cast<float*>(smem_raw + C::X_REGION_B + 2 * C::Z_BUF_B);     // [co][ACC_LD]
    float* scr = acc + CP * C::ACC_LD;                                        // [co][SCR_LD]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + C::X_REGION_B + 2 * C::Z_BUF_B + C::ACC_B + C::SCR_B);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::N_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool pr = (prof != nullptr) && (blockIdx.x == 0);
    long long pw0 = 0, pw1 = 0, pwk = 0, pws = 0;

    for (int i = tid; i < CP * C::ACC_LD; i += C::NTHREADS) acc[i] = 0.f;
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), NSTG); mbar_init(smem_u32(&bars[1]), NSTG);        // a_full: stagers
        mbar_init(smem_u32(&bars[2]), nblk + 6); mbar_init(smem_u32(&bars[3]), nblk + 6);    // a_empty: issuers + producer warps
        for (int b = 0; b < MAXBLK; ++b) {
            mbar_init(smem_u32(&bars[4 + b]), 1);                                         // t_full[b]
            mbar_init(smem_u32(&bars[4 + MAXBLK + b]), 3);                                // t_empty[b]
        }
        for (int s = 0; s < ASLOTS; ++s) {
            mbar_init(smem_u32(&bars[4 + 2 * MAXBLK + s]), 3);                            // s_full[s]: producer warps
            mbar_init(smem_u32(&bars[4 + 2 * MAXBLK + ASLOTS + s]), nblk);                // s_free[s]: one commit per issuer
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[2]);
    const uint32_t t_full = smem_u32(&bars[4]), t_empty = smem_u32(&bars[4 + MAXBLK]);
    const uint32_t s_full = smem_u32(&bars[4 + 2 * MAXBLK]), s_free = smem_u32(&bars[4 + 2 * MAXBLK + ASLOTS]);
    const uint32_t a_tmem = tmem_base + (uint32_t)C::DCOLS;                   // the A ring starts behind the accumulators

    const int ncols = (total_cols - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nsteps = ncols * d.Do;
    const int ndrains = (nsteps + drain_every - 1) / drain_every;

    if (warp >= W_STG0) {
        // =================================================================== staging warps
        const int st = tid - W_STG0 * 32;
        const bool vec_i = (d.ldi % 4 == 0), vec_o = (d.ldo % 4 == 0);
        const bool sc4_i = (d.Ci % 4 == 0), sc4_o = (d.Co % 4 == 0);
        int it = 0, pc = 0;
        for (int col = blockIdx.x; col < total_cols; col += gridDim.x) {
            int t = col;
            const int tw = t % tiles_w; t /= tiles_w;
            const int th_ = t % tiles_h;
            const int n = t / tiles_h;
            const int ow0 = tw * TWW, oh0 = th_ * THW;
            const int ih0 = oh0 - d.ph, iw0 = ow0 - d.pw;
            const int g = n / nPerG;
            const float* xn = X + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;
            const float* zn = dZ + (int64_t)n * d.Do * d.Ho * d.Wo * d.ldo;
            for (int od = 0; od < d.Do; ++od, ++it) {
                const int buf = it & 1, use = it >> 1;
                long long c0 = pr ? clock64() : 0;
                mbar_wait(a_empty + 8 * buf, (use & 1) ^ 1);       // step it-2: MMAs done, dZ tile consumed
                if (NSLOT < 6 && od == 0 && it > 0)                 // short ring: the three planes of a new column overwrite
                    mbar_wait(a_empty + 8 * (buf ^ 1), ((it - 1) >> 1) & 1);   // planes step it-1 still reads
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                const int np = (od == 0) ? 3 : 1;
                const int gd0 = od - d.pd + (3 - np);
                const int nx_items = np * C::XP_ITEMS, n_items = nx_items + C::NZ_ITEMS;
                float* zb = z_reg + (size_t)buf * (C::Z_BUF_B / 4);
#pragma unroll 1
                for (int base = 0; base < n_items; base += NSTG * PER_R) {
                    float4 ra[PER_R], rb[PER_R];
                    int dsto[PER_R];                               // X: byte offset of term 0; dZ: float offset in the tile
                    int meta[PER_R];                               // bits 0..1 group, bit 2 inside the volume, bit 3 dZ item
#pragma unroll
                    for (int u = 0; u < PER_R; ++u) {
                        const int item = base + st + u * NSTG;
                        ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        rb[u] = ra[u];
                        dsto[u] = -1;
                        meta[u] = 0;
                        if (item < nx_items) {
                            const int grp = item % NG;
                            int r = item / NG;
                            const int wx = r % XW; r /= XW;
                            const int hy = r % XH;
                            const int p = r / XH;
                            const int slot = (pc + p) % NSLOT;
                            dsto[u] = ((slot * XH + hy) * NG + grp) * RS + wx * 16;
                            meta[u] = grp;
                            const int gd = gd0 + p, gh = ih0 + hy, gw = iw0 + wx;
                            const int c = grp * 8;
                            if (gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi && c < d.Ci) {
                                meta[u] |= 4;
                                const float* pp = xn + (((int64_t)gd * d.Hi + gh) * d.Wi + gw) * d.ldi + c;
                                if (vec_i && c + 8 <= d.Ci) {
                                    ra[u] = *reinterpret_cast<const float4*>(pp);
                                    rb[u] = *reinterpret_cast<const float4*>(pp + 4);
                                } else {
                                    float e[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) e[j] = (c + j < d.Ci) ? pp[j] : 0.f;
                                    ra[u] = make_float4(e[0], e[1], e[2], e[3]);
                                    rb[u] = make_float4(e[4], e[5], e[6], e[7]);
                                }
                            }
                        } else if (item < n_items) {
                            const int j = item - nx_items;
                            const int grp = j % NG, v = j / NG;                // voxel of the output tile, row-major
                            dsto[u] = v * CP + grp * 8;
                            meta[u] = grp | 8;
                            const int gh = oh0 + v / TWW, gw = ow0 + v % TWW;
                            const int c = grp * 8;
                            if (gh < d.Ho && gw < d.Wo && c < d.Co) {
                                meta[u] |= 4;
                                const float* pp = zn + (((int64_t)od * d.Ho + gh) * d.Wo + gw) * d.ldo + c;
                                if (vec_o && c + 8 <= d.Co) {
                                    ra[u] = *reinterpret_cast<const float4*>(pp);
                                    rb[u] = *reinterpret_cast<const float4*>(pp + 4);
                                } else {
                                    float e[8];
#pragma unroll
                                    for (int jj = 0; jj < 8; ++jj) e[jj] = (c + jj < d.Co) ? pp[jj] : 0.f;
                                    ra[u] = make_float4(e[0], e[1], e[2], e[3]);
                                    rb[u] = make_float4(e[4], e[5], e[6], e[7]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PER_R; ++u) {
                        if (dsto[u] < 0) continue;
                        float v[8] = {ra[u].x, ra[u].y, ra[u].z, ra[u].w, rb[u].x, rb[u].y, rb[u].z, rb[u].w};
                        const bool isz = (meta[u] & 8) != 0;
                        const int c = (meta[u] & 3) * 8;
                        if (meta[u] & 4) {
                            const float* scp = isz ? o_scale : i_scale;
                            const float* shp = isz ? o_shift : i_shift;
                            const int CC = isz ? d.Co : d.Ci;
                            if (scp) {
                                if ((isz ? sc4_o : sc4_i) && c + 8 <= CC) {
                                    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scp + (int64_t)g * CC + c));
                                    const float4 s1 = __ldg(reinterpret_cast<const float4*>(scp + (int64_t)g * CC + c + 4));
                                    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shp + (int64_t)g * CC + c));
                                    const float4 h1 = __ldg(reinterpret_cast<const float4*>(shp + (int64_t)g * CC + c + 4));
                                    v[0] = fmaf(v[0], s0.x, h0.x); v[1] = fmaf(v[1], s0.y, h0.y);
                                    v[2] = fmaf(v[2], s0.z, h0.z); v[3] = fmaf(v[3], s0.w, h0.w);
                                    v[4] = fmaf(v[4], s1.x, h1.x); v[5] = fmaf(v[5], s1.y, h1.y);
                                    v[6] = fmaf(v[6], s1.z, h1.z); v[7] = fmaf(v[7], s1.w, h1.w);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 8; ++j)
                                        if (c + j < CC) v[j] = fmaf(v[j], scp[(int64_t)g * CC + c + j], shp[(int64_t)g * CC + c + j]);
                                }
                            }
                        }
                        if (isz) {                                 // raw fp32 tile for the TMEM producers
                            float4* zp = reinterpret_cast<float4*>(zb + dsto[u]);
                            zp[0] = make_float4(v[0], v[1], v[2], v[3]);
                            zp[1] = make_float4(v[4], v[5], v[6], v[7]);
                        } else {
                            uint4 o[3];
                            split8_trunc3(v, o);
#pragma unroll
                            for (int s2 = 0; s2 < 3; ++s2) *reinterpret_cast<uint4*>(x_reg + s2 * C::X_TERM_B + dsto[u]) = o[s2];
                        }
                    }
                }
                pc += np;
                fence_async_smem();
                mbar_arrive(a_full + 8 * buf);
                if (pr) pwk += clock64() - c1;
            }
        }
        if (pr && st == 0) { prof[4] = pw0; prof[5] = pwk; }
    } else if (warp == 3 || warp == 7 || warp >= 11) {
        // =================================================================== MMA issue: issuer i owns block b0 + i
        const int bi = (warp == 3) ? 0 : (warp == 7) ? 1 : warp - 9;
        if (lane == 0 && bi < nblk) {
            const int b = b0 + bi, kd = b / 3, kw = b % 3;
            const uint32_t x_base = smem_u32(x_reg);
            const uint32_t dcol = tmem_base + (uint32_t)(bi * BCOLS);
            constexpr uint32_t IDESC = idesc_ts(BCOLS);
            bool fresh = true;
            int drains = 0, pc = 0, gk = 0;                       // gk: K steps issued so far (A ring position)
            for (int it = 0; it < nsteps; ++it) {
                const int buf = it & 1, use = it >> 1;
                pc += (it % d.Do == 0) ? 3 : 1;
                long long c0 = pr ? clock64() : 0;
                mbar_wait(a_full + 8 * buf, use & 1);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                if (fresh && drains > 0) mbar_wait(t_empty + 8 * bi, (drains - 1) & 1);
                long long c2 = pr ? clock64() : 0;
                pw1 += c2 - c1;
                tc_fence_after();
                const int slot = (pc - 3 + kd) % NSLOT;
                const uint64_t db0 = umma_desc(x_base + (uint32_t)(slot * C::X_PLANE_B + kw * 16), 128, RS);
#pragma unroll 1
                for (int ks = 0; ks < KSTEPS; ++ks, ++gk) {
                    const int as = gk % ASLOTS;
                    long long cs = pr ? clock64() : 0;
                    mbar_wait(s_full + 8 * as, (gk / ASLOTS) & 1);
                    if (pr) pws += clock64() - cs;
                    tc_fence_after();
                    const int r = ks / (TWW / 16), kc = ks % (TWW / 16);
                    const uint64_t db = db0 + (uint64_t)((r * C::X_ROW_B + kc * 256) >> 4);
                    const uint32_t at = a_tmem + (uint32_t)(as * 8);
#pragma unroll
                    for (int tx = 2; tx >= 0; --tx) {
                        umma_bf16_ts(dcol, at, db + (uint64_t)((tx * C::X_TERM_B) >> 4), IDESC, fresh ? 0u : 1u);
                        fresh = false;
                    }
                    umma_commit(s_free + 8 * as);
                }
                umma_commit(a_empty + 8 * buf);
                if ((it + 1) % drain_every == 0 || it == nsteps - 1) {
                    umma_commit(t_full + 8 * bi);
                    fresh = true;
                    ++drains;
                }
                if (pr) pwk += clock64() - c2;
            }
            if (pr && bi == 0) { prof[0] = pw0; prof[1] = pw1; prof[2] = pwk; prof[3] = nsteps; prof[8] = pws; }
        }
    } else if (warp >= 4) {
        // =================================================================== A producers: warps 4-6 / 8-10 write term warp % 4
        // of the even / odd K steps, lane = co
        const int term = warp & 3, pset = (warp >> 2) - 1;
        {
            const uint32_t tq = a_tmem + ((uint32_t)(term * 32) << 16);
            for (int it = 0; it < nsteps; ++it) {
                const int buf = it & 1, use = it >> 1;
                mbar_wait(a_full + 8 * buf, use & 1);
                const float* zb = z_reg + (size_t)buf * (C::Z_BUF_B / 4) + (lane < CP ? lane : 0);
#pragma unroll 1
                for (int ks = pset; ks < KSTEPS; ks += 2) {
                    const int gk = it * KSTEPS + ks;
                    const int as = gk % ASLOTS;
                    long long cs = pr ? clock64() : 0;
                    if (gk >= ASLOTS) mbar_wait(s_free + 8 * as, ((gk / ASLOTS) - 1) & 1);
                    long long ce = pr ? clock64() : 0;
                    pw0 += ce - cs;
                    tc_fence_after();
                    uint32_t h[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float v = (lane < CP) ? zb[(ks * 16 + k) * CP] : 0.f;
                        const uint32_t t1 = __float_as_uint(v) & 0xffff0000u;
                        const float r1 = v - __uint_as_float(t1);
                        const uint32_t t2 = __float_as_uint(r1) & 0xffff0000u;
                        const float r2 = r1 - __uint_as_float(t2);
                        h[k] = (term == 0) ? t1 : (term == 1) ? t2 : __float_as_uint(r2);
                    }
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) pk[j] = __byte_perm(h[2 * j], h[2 * j + 1], 0x7632);   // low half = even voxel
                    tmem_st8(tq + (uint32_t)(as * 8), pk);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_full + 8 * as);
                    if (pr) pwk += clock64() - ce;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(a_empty + 8 * buf);     // this warp no longer reads the dZ tile of the step
            }
            if (pr && warp == 4 && lane == 0) { prof[9] = pw0; prof[10] = pwk; }
        }
    } else if (warp < 3) {
        // =================================================================== drain: warp t = y term t, lane = co
        const int co = lane < CP ? lane : CP - 1;
        float* srow = scr + co * C::SCR_LD;
        float* arow = acc + co * C::ACC_LD;
        for (int dr = 0; dr < ndrains; ++dr) {
#pragma unroll 1
            for (int bi = 0; bi < nblk; ++bi) {
                long long c0 = pr ? clock64() : 0;
                mbar_wait(t_full + 8 * bi, dr & 1);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                tc_fence_after();
                float v[BCOLS];
                tmem_ld_cols<BCOLS>(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(bi * BCOLS), v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_empty + 8 * bi);
                // (term 2 + term 1) + term 0, then one addition to the running fp32 sum
                if (warp == 2 && lane < CP) {
#pragma unroll
                    for (int j4 = 0; j4 < BCOLS / 4; ++j4)
                        reinterpret_cast<float4*>(srow)[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                }
                drain_bar();
                if (warp == 1 && lane < CP) {
#pragma unroll
                    for (int j4 = 0; j4 < BCOLS / 4; ++j4) {
                        float4 s = reinterpret_cast<float4*>(srow)[j4];
                        s.x += v[4 * j4]; s.y += v[4 * j4 + 1]; s.z += v[4 * j4 + 2]; s.w += v[4 * j4 + 3];
                        reinterpret_cast<float4*>(srow)[j4] = s;
                    }
                }
                drain_bar();
                if (warp == 0 && lane < CP) {
                    float4* ap = reinterpret_cast<float4*>(arow + bi * BCOLS);
#pragma unroll
                    for (int j4 = 0; j4 < BCOLS / 4; ++j4) {
                        const float4 s = reinterpret_cast<float4*>(srow)[j4];
                        float4 a = ap[j4];
                        a.x += s.x + v[4 * j4]; a.y += s.y + v[4 * j4 + 1]; a.z += s.z + v[4 * j4 + 2]; a.w += s.w + v[4 * j4 + 3];
                        ap[j4] = a;
                    }
                }
                drain_bar();
                if (pr) pwk += clock64() - c1;
            }
        }
        if (pr && tid == 0) { prof[6] = pw0; prof[7] = pwk; }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
    const int wn = d.Co * d.Ci * 27;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = tid; i < wn; i += C::NTHREADS) {
        const int tap = i % 27, ci = (i / 27) % d.Ci, co = i / (27 * d.Ci);
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        const int bi = kd * 3 + kw - b0;
        if (bi >= 0 && bi < nblk) wsp[i] = acc[co * C::ACC_LD + bi * BCOLS + kh * CP + ci];
    }
}

}  // namespace sp_wtca

template <int NG>
static inline int sp_tca_wgrad_launch_t(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                        const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                        cudaStream_t st, long long* prof, int drain_every) {
    using namespace sp_wtca;
    using C = Cfg<NG>;
    const sp_wtc::WtcPlan p = sp_wtc::plan(d);
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(wgrad3_tca_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        attr = true;
    }
    const int npass = (9 + C::MAXBLK - 1) / C::MAXBLK;
    for (int pass = 0; pass < npass; ++pass) {
        const int b0 = pass * C::MAXBLK, nblk = (9 - b0 < C::MAXBLK) ? 9 - b0 : C::MAXBLK;
        wgrad3_tca_kernel<NG><<<p.grid, C::NTHREADS, C::SMEM, st>>>(*d, nPerG, p.tiles_w, p.tiles_h, (int)p.total, drain_every, b0, nblk,
                                                                    iside, i_scale, i_shift, oside, o_scale, o_shift, ws,
                                                                    pass == 0 ? prof : nullptr);
        SP_LAUNCH_OK("wgrad3_tca_kernel");
    }
    const int64_t wn = (int64_t)d->Co * d->Ci * 27;
    wgrad_reduce_kernel<<<(int)((wn + 255) / 256), 256, 0, st>>>(ws, p.grid, wn, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}

// SP_WGRAD_TCA=0 keeps the first-generation kernels (A operand from shared memory)
static inline bool sp_tca_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SP_WGRAD_TCA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

static inline int sp_tca_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                      const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                      cudaStream_t st, long long* prof = nullptr, int drain_every = 2) {
    if (d->Ci <= 16 && d->Co <= 16)
        return sp_tca_wgrad_launch_t<2>(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, ws, st, prof, drain_every);
    return sp_tca_wgrad_launch_t<3>(d, nPerG, iside, i_scale, i_shift, oside, o_scale, o_shift, dw, beta, ws, st, prof, drain_every);
}
