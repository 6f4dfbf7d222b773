// sp_wgrad_tc24.cuh — tcgen05 / TMEM weight gradient of the 3x3x3 stride-1 layers with 17..24 channels on either side
// (Cae3D.py:52,55,186,189,197,200: the 24-channel level of the CAE, 30 % of its FFMA-tier wgrad time at 22-29 TFLOP/s).
// Same formulation as sp_wgrad_tc.cuh (voxels = contraction index, MN-major operands, dZ's three bf16 terms stacked along
// M, a CTA walks a 4 x 32 column along the depth axis through a ring of input planes) with three 8-channel groups per side:
//   * A (dZ): M = 128, rows 32 t + co (M group 4 t + g: term t sits in TMEM lane quarter t), 72 useful rows;
//   * B (X'): N = (kh, ci) = 72 columns per MMA, N-group stride RS inside an input row [group][w], rows 3 RS apart;
//   * nine (kd, kw) accumulator blocks x 72 columns = 648 TMEM columns do not fit: TWO passes over the volume, blocks 0..4
//     (360 columns) then blocks 5..8 (288), each a launch of this kernel writing its taps of the per-CTA partial;
//   * the accumulators are drained every `drain_every` steps; the three y terms are folded smallest first through a small
//     shared scratch tile (term 2 -> + term 1 -> + term 0) and added ONCE to the fp32 sums acc[co][block col].
// Even with nine bf16 products per fp32 MAC the tensor pipe beats the FFMA pipe several times over at these widths.
#pragma once
#include "sp_wgrad_tc.cuh"

namespace sp_wtc24 {

using namespace sp_tc;
using sp_tc2::mbar_arrive;
using sp_tc2::split8_trunc3;

constexpr int TWW = 32, THW = 4;
constexpr int XW = TWW + 2, XH = THW + 2;
constexpr int NG = 3;                              // 8-channel groups per side
constexpr int RS = XW * 16;                        // bytes per (row, group) = N-group stride
constexpr int X_ROW_B = NG * RS;
constexpr int X_PLANE_B = XH * X_ROW_B;            // 9792
constexpr int NSLOT = 4;                           // ring of input planes (a column start waits for the previous step)
constexpr int X_TERM_B = NSLOT * X_PLANE_B;
constexpr int X_REGION_B = 3 * X_TERM_B;           // 117504
constexpr int PS = TWW * THW * 16;                 // bytes per (term, group) plane of the dZ tile = M-group stride
constexpr int A_BUF_B = 12 * PS;                   // [term][4 group slots, 3 used]: term t starts at M group 4 t
constexpr int A_REGION_B = 2 * A_BUF_B;            // 49152; the 16 M groups read from buffer 1 end inside the X region
constexpr int BCOLS = 72, MAXBLK = 5;
constexpr int ACC_LD = MAXBLK * BCOLS + 4;         // 364 = 4 * 91
constexpr int ACC_B = 24 * ACC_LD * 4;
constexpr int SCR_LD = BCOLS + 4;                  // 76 = 4 * 19
constexpr int SCR_B = 24 * SCR_LD * 4;
constexpr int W_EPI = 4, W_MMA = MAXBLK, W_STG = 10;
constexpr int NSTG = W_STG * 32;
constexpr int NTHREADS_W = (W_EPI + W_MMA + W_STG) * 32;      // 544
constexpr int XP_ITEMS = XH * XW * NG, NZ_ITEMS = TWW * THW * NG;
constexpr int PER_R = 4;
constexpr int N_BARS = 2 + 2 + MAXBLK + MAXBLK;
constexpr size_t SMEM_W = (size_t)A_REGION_B + X_REGION_B + ACC_B + SCR_B + N_BARS * 8 + 16;
static_assert(SMEM_W <= 227 * 1024, "wgrad tc24: shared memory");
static_assert(A_BUF_B + 16 * PS <= A_REGION_B + X_REGION_B, "M groups of buffer 1 stay inside the A + X regions");

// 72 consecutive TMEM columns = x64 + x8, one wait
__device__ __forceinline__ void tmem_ld72(uint32_t taddr, float* v) {
    uint32_t r[72];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
        "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
          "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
          "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[64]), "=r"(r[65]), "=r"(r[66]), "=r"(r[67]), "=r"(r[68]), "=r"(r[69]), "=r"(r[70]), "=r"(r[71])
                 : "r"(taddr + 64) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 72; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void drain_bar() { asm volatile("bar.sync 1, 96;" ::: "memory"); }

// blocks [b0, b0 + nblk) of the nine (kd, kw) accumulator blocks (b = kd * 3 + kw).
// MROWS = 64: at most 16 output channels (24 -> 16, Cae3D.py:200): two O-side groups per term, M groups 2 t + g, rows 16 t + co
// = TMEM lanes 32 t + co as well, and half the A-operand fetch of the M = 128 form.
template <int MROWS>
__global__ void __launch_bounds__(NTHREADS_W, 1)
wgrad3_tc24_kernel(SpConvDesc d, int nPerG, int tiles_w, int tiles_h, int total_cols, int drain_every, int b0, int nblk,
                   const float* __restrict__ X, const float* __restrict__ i_scale, const float* __restrict__ i_shift,
                   const float* __restrict__ dZ, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                   float* __restrict__ ws, long long* __restrict__ prof, int nt) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* a_reg = smem_raw;                                          // [buf][term][4][voxel] x 16 B
    unsigned char* x_reg = smem_raw + A_REGION_B;                             // [term][slot][row][group][w] x 16 B
    float* acc = reinterpret_cast<float*>(smem_raw + A_REGION_B + X_REGION_B);    // [co][ACC_LD]
    float* scr = acc + 24 * ACC_LD;                                           // [co][SCR_LD]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + A_REGION_B + X_REGION_B + ACC_B + SCR_B);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool pr = (prof != nullptr) && (blockIdx.x == 0);
    long long pw0 = 0, pw1 = 0, pwk = 0;

    for (int i = tid; i < 24 * ACC_LD; i += NTHREADS_W) acc[i] = 0.f;
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), NSTG); mbar_init(smem_u32(&bars[1]), NSTG);        // a_full
        mbar_init(smem_u32(&bars[2]), nblk); mbar_init(smem_u32(&bars[3]), nblk);        // a_empty: one commit per issuer
        for (int b = 0; b < MAXBLK; ++b) {
            mbar_init(smem_u32(&bars[4 + b]), 1);                                         // t_full[b]
            mbar_init(smem_u32(&bars[4 + MAXBLK + b]), 3);                                // t_empty[b]: drain warps 0..2
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

    const int ncols = (total_cols - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nsteps = ncols * d.Do;
    const int ndrains = (nsteps + drain_every - 1) / drain_every;

    if (warp >= W_EPI + W_MMA) {
        // =================================================================== staging warps
        const int st = tid - (W_EPI + W_MMA) * 32;
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
                mbar_wait(a_empty + 8 * buf, (use & 1) ^ 1);       // the MMAs of step it-2 are done
                if (od == 0 && it > 0)                              // four slots: the three planes of a new column overwrite
                    mbar_wait(a_empty + 8 * (buf ^ 1), ((it - 1) >> 1) & 1);   // planes step it-1 still reads
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                const int np = (od == 0) ? 3 : 1;
                const int gd0 = od - d.pd + (3 - np);
                const int nx_items = np * XP_ITEMS, n_items = nx_items + NZ_ITEMS;
#pragma unroll 1
                for (int base = 0; base < n_items; base += NSTG * PER_R) {
                    float4 ra[PER_R], rb[PER_R];
                    int dsto[PER_R];                               // byte offset of term 0 (-1: no item)
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
                            if (MROWS == 64 && grp == 2) continue;             // no third O-side group in the M = 64 layout
                            dsto[u] = buf * A_BUF_B + grp * PS + v * 16;
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
                            const int C = isz ? d.Co : d.Ci;
                            if (scp) {
                                if ((isz ? sc4_o : sc4_i) && c + 8 <= C) {
                                    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scp + (int64_t)g * C + c));
                                    const float4 s1 = __ldg(reinterpret_cast<const float4*>(scp + (int64_t)g * C + c + 4));
                                    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shp + (int64_t)g * C + c));
                                    const float4 h1 = __ldg(reinterpret_cast<const float4*>(shp + (int64_t)g * C + c + 4));
                                    v[0] = fmaf(v[0], s0.x, h0.x); v[1] = fmaf(v[1], s0.y, h0.y);
                                    v[2] = fmaf(v[2], s0.z, h0.z); v[3] = fmaf(v[3], s0.w, h0.w);
                                    v[4] = fmaf(v[4], s1.x, h1.x); v[5] = fmaf(v[5], s1.y, h1.y);
                                    v[6] = fmaf(v[6], s1.z, h1.z); v[7] = fmaf(v[7], s1.w, h1.w);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 8; ++j)
                                        if (c + j < C) v[j] = fmaf(v[j], scp[(int64_t)g * C + c + j], shp[(int64_t)g * C + c + j]);
                                }
                            }
                        }
                        uint4 o[3];
                        if (nt == 1) {      // bf16 mode: ONE bf16 term (round to nearest); the other two term planes are zero
                            o[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                            o[1] = o[2] = make_uint4(0u, 0u, 0u, 0u);
                        } else {
                            split8_trunc3(v, o);
                        }
                        unsigned char* dstp = (isz ? a_reg : x_reg) + dsto[u];
                        const int tstride = isz ? (MROWS == 128 ? 4 : 2) * PS : X_TERM_B;
#pragma unroll
                        for (int s2 = 0; s2 < 3; ++s2) *reinterpret_cast<uint4*>(dstp + s2 * tstride) = o[s2];
                    }
                }
                pc += np;
                fence_async_smem();
                mbar_arrive(a_full + 8 * buf);
                if (pr) pwk += clock64() - c1;
            }
        }
        if (pr && st == 0) { prof[4] = pw0; prof[5] = pwk; }
    } else if (warp >= W_EPI) {
        // =================================================================== MMA issue: warp 4 + i owns block b0 + i
        // (warp-converged issue: the whole warp runs the loop with warp-uniform values, one elected lane issues — umma_bf16_elect)
        const int bi = sp_tc::warp_uniform(warp - W_EPI);
        if (bi < nblk) {
            const int b = b0 + bi, kd = b / 3, kw = b % 3;
            const uint32_t a_base = smem_u32(a_reg), x_base = smem_u32(x_reg);
            const uint32_t dcol = tmem_base + (uint32_t)(bi * BCOLS);
            constexpr uint32_t IDESC = sp_wtc::idesc_mn(MROWS, BCOLS);
            bool fresh = true;
            int drains = 0, pc = 0;
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
                const uint64_t da0 = umma_desc(a_base + (uint32_t)(buf * A_BUF_B), 128, PS);
                const uint64_t db0 = umma_desc(x_base + (uint32_t)(slot * X_PLANE_B + kw * 16), 128, RS);
#pragma unroll 1
                for (int r = 0; r < THW; ++r) {
#pragma unroll
                    for (int ks = 0; ks < TWW / 16; ++ks) {
                        const uint64_t da = da0 + (uint64_t)(r * TWW + ks * 16);
                        const uint64_t db = db0 + (uint64_t)((r * X_ROW_B + ks * 256) >> 4);
                        for (int tx = nt - 1; tx >= 0; --tx) {
                            umma_bf16_elect(dcol, da, db + (uint64_t)((tx * X_TERM_B) >> 4), IDESC, fresh ? 0u : 1u);
                            fresh = false;
                        }
                    }
                }
                umma_commit_elect(a_empty + 8 * buf);
                if ((it + 1) % drain_every == 0 || it == nsteps - 1) {
                    umma_commit_elect(t_full + 8 * bi);
                    fresh = true;
                    ++drains;
                }
                if (pr) pwk += clock64() - c2;
            }
            if (pr && bi == 0 && lane == 0) { prof[0] = pw0; prof[1] = pw1; prof[2] = pwk; prof[3] = nsteps; }
        }
    } else if (warp < 3) {
        // =================================================================== drain: warp t = y term t, lane = co
        const int co = lane < 24 ? lane : 23;
        float* srow = scr + co * SCR_LD;
        float* arow = acc + co * ACC_LD;
        for (int dr = 0; dr < ndrains; ++dr) {
#pragma unroll 1
            for (int bi = 0; bi < nblk; ++bi) {
                long long c0 = pr ? clock64() : 0;
                mbar_wait(t_full + 8 * bi, dr & 1);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                tc_fence_after();
                float v[BCOLS];
                tmem_ld72(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(bi * BCOLS), v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_empty + 8 * bi);
                // (term 2 + term 1) + term 0, then one addition to the running fp32 sum
                if (warp == 2 && lane < 24) {
#pragma unroll
                    for (int j4 = 0; j4 < BCOLS / 4; ++j4)
                        reinterpret_cast<float4*>(srow)[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                }
                drain_bar();
                if (warp == 1 && lane < 24) {
#pragma unroll
                    for (int j4 = 0; j4 < BCOLS / 4; ++j4) {
                        float4 s = reinterpret_cast<float4*>(srow)[j4];
                        s.x += v[4 * j4]; s.y += v[4 * j4 + 1]; s.z += v[4 * j4 + 2]; s.w += v[4 * j4 + 3];
                        reinterpret_cast<float4*>(srow)[j4] = s;
                    }
                }
                drain_bar();
                if (warp == 0 && lane < 24) {
                    float4* ap = reinterpret_cast<float4*>(arow + bi * BCOLS);
#pragma unroll
                    for (int j4 = 0; j4 < BCOLS / 4; ++j4) {
                        const float4 s = reinterpret_cast<float4*>(srow)[j4];
                        float4 a = ap[j4];
                        a.x += s.x + v[4 * j4]; a.y += s.y + v[4 * j4 + 1]; a.z += s.z + v[4 * j4 + 2]; a.w += s.w + v[4 * j4 + 3];
                        ap[j4] = a;
                    }
                }
                drain_bar();                                   // the scratch tile is free for the next block
                if (pr) pwk += clock64() - c1;
            }
        }
        if (pr && tid == 0) { prof[6] = pw0; prof[7] = pwk; }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
    // this pass's taps of the CTA partial, torch layout dW[co][ci][tap]
    const int wn = d.Co * d.Ci * 27;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = tid; i < wn; i += NTHREADS_W) {
        const int tap = i % 27, ci = (i / 27) % d.Ci, co = i / (27 * d.Ci);
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        const int bi = kd * 3 + kw - b0;
        if (bi >= 0 && bi < nblk) wsp[i] = acc[co * ACC_LD + bi * BCOLS + kh * 24 + ci];
    }
}

}  // namespace sp_wtc24

static inline bool sp_tc24_wgrad_supported(const SpConvDesc* d) {
    if (d->k != 3 || d->s != 1 || sp_tc_terms() == 0 || sp_tc_wgrad_disabled()) return false;
    if (d->Ci <= 8 || d->Ci > 24 || d->Co <= 8 || d->Co > 24 || (d->Ci <= 16 && d->Co <= 16)) return false;
    if (d->pd > 2 || d->ph > 2 || d->pw > 2) return false;
    const sp_wtc::WtcPlan p = sp_wtc::plan(d);
    return p.total >= 16 && p.total < (1LL << 31) && d->Wo >= 24 && d->Do >= 8;
}

static inline size_t sp_tc24_wgrad_workspace_bytes(const SpConvDesc* d) {
    if (!sp_tc24_wgrad_supported(d)) return 0;
    return (size_t)sp_wtc::plan(d).grid * d->Co * d->Ci * 27 * sizeof(float);
}

static inline int sp_tc24_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                       const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                       cudaStream_t st, long long* prof = nullptr, int drain_every = 2) {
    using namespace sp_wtc24;
    const sp_wtc::WtcPlan p = sp_wtc::plan(d);
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(wgrad3_tc24_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_W));
        SP_CUDA(cudaFuncSetAttribute(wgrad3_tc24_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_W));
        attr = true;
    }
    for (int pass = 0; pass < 2; ++pass) {
        const int b0 = pass == 0 ? 0 : MAXBLK, nblk = pass == 0 ? MAXBLK : 9 - MAXBLK;
        if (d->Co <= 16)
            wgrad3_tc24_kernel<64><<<p.grid, NTHREADS_W, SMEM_W, st>>>(*d, nPerG, p.tiles_w, p.tiles_h, (int)p.total, drain_every, b0, nblk,
                                                                       iside, i_scale, i_shift, oside, o_scale, o_shift, ws,
                                                                       pass == 0 ? prof : nullptr, sp_tc_terms() == 1 ? 1 : 3);
        else
            wgrad3_tc24_kernel<128><<<p.grid, NTHREADS_W, SMEM_W, st>>>(*d, nPerG, p.tiles_w, p.tiles_h, (int)p.total, drain_every, b0, nblk,
                                                                        iside, i_scale, i_shift, oside, o_scale, o_shift, ws,
                                                                        pass == 0 ? prof : nullptr, sp_tc_terms() == 1 ? 1 : 3);
        SP_LAUNCH_OK("wgrad3_tc24_kernel");
    }
    const int64_t wn = (int64_t)d->Co * d->Ci * 27;
    wgrad_reduce_kernel<<<(int)((wn + 255) / 256), 256, 0, st>>>(ws, p.grid, wn, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}
