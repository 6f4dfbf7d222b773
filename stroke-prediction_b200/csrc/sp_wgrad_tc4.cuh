// sp_wgrad_tc4.cuh — second-generation tcgen05 / TMEM weight gradient of the 3x3x3 stride-1 layers: 16 x 16 channels per slice
// pair (Cae3D.py:44,208,211; Unet3D.py:22 — the three largest kernels of the round-1 CAE training step), wider layers (up to
// 96 -> 64: Cae3D.py:52,55,186-200; Unet3D.py:19,22) as several slice pairs walked by ONE launch, thin inputs (2..8 channels).
//
//   dW[co][ci][kd][kh][kw] = sum_{n,od,oh,ow} dZ[n,od,oh,ow,co] * X'[n, od-pd+kd, oh-ph+kh, ow-pw+kw, ci]
//
// The first generation (sp_wgrad_tc.cuh) issues 27 MMAs of M = 64, N = 48 per 16 voxels (one per (kd, kw) block and X'
// term; 9 bf16 products per fp32 MAC, a quarter of the 128 x N datapath in use) and is bound by the shared-memory fetch of
// those 216 small MMAs per step.  Here ONE MMA of M = 128 (96 rows used), N = 96 per (16 voxels, kd) forms every product:
//
//   K          = 16 consecutive voxels u of one INPUT row (the CTA's tile is a 32-wide range of X columns, no w halo);
//   A (M side) = dZ, rows (term_y, kw, co): the kw tap is a COPY of the dZ row shifted by kw voxels (ow = u + pw - kw;
//                the staging thread that holds a dZ value stores it three times), the two bf16 terms of every value are
//                stacked as well: 2 x 3 x 16 = 96 rows, M group stride PS;
//   B (N side) = X' (BatchNorm applied, padding = zeros), columns (kh, term_x, ci): the three input rows oh+kh of one
//                output row and both terms lie one uniform N-group stride RS apart ([row][term][half][u] planes): N = 96;
//   kd         = slot of the depth ring in the descriptor's start address -> three accumulator blocks of 96 TMEM columns,
//                one issuing warp each (MMAs of one thread retire in order, MMAs of different warps overlap).
// Both operands are MN-major (channels contiguous, 8 voxels x 16 B core matrices), as in the first generation.
//
// Arithmetic: every fp32 value v is staged as TWO bf16 terms, both rounded to nearest: t1 = rn(v), t2 = rn(v - t1), so
// |v - t1 - t2| <= 2^-18 |v|; all four products y_i x_j are formed.  The rounding errors are unbiased and independent from
// voxel to voxel, so they average out over the >= 10^5 voxels of a sum: measured rel-L2 of dW against an fp64 sum is the same
// as with the exact three-term split of the first generation (the TMEM accumulator's truncation dominates both;
// profiles/r02_wgrad_tc4_probe.log) while the tensor pipe does 4 bf16 MACs per fp32 MAC instead of 12 issued.
// The accumulators are drained every `drain_every` steps (8 accumulations per step) into fp32 round-to-nearest sums in
// shared memory, the x terms folded on the way; the y terms are folded at the end, small first.
//
// Persistent CTA per SM, 28 warps: 0-2 drain TMEM lane quarters 0-2, 4-6 issue the MMAs of block kd (warp-converged, one
// elected lane: see umma_bf16_elect), 7-27 stage — three groups of seven warps that take every third step (group = A
// buffer), so the global-load latency of steps i+1, i+2 overlaps the conversion of step i; the loads of a step are issued
// BEFORE its buffer is waited for.  Inside a group four warps stage the X' plane(s) and three the dZ tile, every thread with
// a fixed (row set, channel half, column), so a step costs a few pointer increments.  A CTA walks a column (n, 4 output
// rows, 32 input columns) along the depth axis through a ring of seven X' planes; the first two planes of the NEXT column
// are staged during the last two steps of the current one, so every step but the CTA's first stages one or two planes.
// Work items = (slice pair, tile column), dealt out round-robin over the CTAs; the drain warps flush the accumulators to
// ws[pair][cta] as [co 16][ci 16][27] at every pair change, wgrad_reduce_pairs_kernel folds them in CTA order (deterministic).
#pragma once
#include "sp_wgrad_tc.cuh"

// run-time options (sp_set_wgrad_tc_options): generation 1 keeps sp_wgrad_tc.cuh, max_ctas > 0 caps the persistent grid;
// initial values from SP_DISABLE_TC4_WGRAD=1 / SP_WTC4_GRID=n
static inline int& sp_wtc4_grid_cap_ref() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SP_WTC4_GRID");
        v = e ? atoi(e) : 0;
        if (v < 0) v = 0;
    }
    return v;
}
static inline int& sp_wtc4_generation_ref() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SP_DISABLE_TC4_WGRAD");
        v = (e && e[0] == '1') ? 1 : 2;
    }
    return v;
}

namespace sp_wtc4 {

using namespace sp_tc;
using sp_tc2::mbar_arrive;
using sp_wtc::idesc_mn;

constexpr int TU = 32, THW = 4;                    // tile: 32 input columns x 4 output rows
constexpr int XH = THW + 2;                        // input rows of a tile
constexpr int ZW = TU + 2;                         // dZ columns a tile reads (ow = u + pw - kw)
constexpr int RS = TU * 16;                        // bytes per (row, term, half) of an X' plane = N-group stride (512)
constexpr int X_PLANE_B = XH * 4 * RS;             // [row][term][half][u] = 12288
#ifndef SP_WTC4_NBUF
#define SP_WTC4_NBUF 3
#endif
constexpr int NBUF = SP_WTC4_NBUF;                 // dZ tile buffers = staging groups = steps the staging may run ahead
constexpr int NSLOT = NBUF + 4;                    // depth ring (see the safety argument at the staging schedule)
constexpr int X_REGION_B = NSLOT * X_PLANE_B;      // 73728
constexpr int PS = THW * TU * 16;                  // bytes per (term, kw, half) plane of the dZ tile = M-group stride (2048)
constexpr int A_BUF_B = 12 * PS;                   // [term][kw][half][row][u] = 24576
constexpr int A_REGION_B = NBUF * A_BUF_B;
constexpr int NBLK = 3, BCOLS = 96;                // accumulator blocks (kd) x columns (kh, term_x, ci)
constexpr int ACC_ROWS = 96;                       // (term_y, kw, co)
constexpr int ACC_LD = 148;                        // floats per row: 144 columns (kd, kh, ci) + pad (4 x odd: conflict-free rows)
constexpr int ACC_B = ACC_ROWS * ACC_LD * 4;       // 56832
constexpr int MAXG = 16;                           // BatchNorm statistics groups whose coefficients fit the shared table
constexpr int MAXSI = 6;                            // 16-channel input slices of one launch (96 channels)
constexpr int COEF_B = MAXG * MAXSI * 32 * 4;
constexpr int WS_SLOT = 16 * 16 * 27;               // floats of one (CTA, slice pair) partial: [co 16][ci 16][27]
constexpr int W_EPI = 4, W_MMA = 3, W_STG_X = 4, W_STG_Z = 3, W_STG_G = W_STG_X + W_STG_Z, NGRP = NBUF;
static_assert(W_STG_X * 32 == 2 * 2 * TU && W_STG_Z * 32 >= 2 * ZW && XH == 6, "staging roles");
constexpr int NTHREADS4 = (W_EPI + W_MMA + NGRP * W_STG_G) * 32;   // 736
constexpr int N_BARS = 2 * NBUF + 2 * NBLK;
constexpr size_t SMEM4 = (size_t)A_REGION_B + X_REGION_B + ACC_B + COEF_B + N_BARS * 8 + 16;
static_assert(SMEM4 <= 227 * 1024, "wgrad tc4: shared memory");
static_assert(A_REGION_B + X_REGION_B >= (NBUF - 1) * A_BUF_B + 16 * PS, "the 16 M groups of the last buffer stay inside shared memory");

// Waits use mbarrier.try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the
// hint expires), so a waiting warp issues a handful of instructions instead of spinning — with plain polling half of all
// issued instructions of this kernel were TRYWAIT / NANOSLEEP / BRA of waiting warps, taken from the issue slots of the
// staging warps on the same schedulers (ncu source page, profiles/r02_wgrad_tc4_notes.md).  Bounded: a wrong descriptor
// must fail the launch (trap), never hang the GPU.
__device__ __forceinline__ void mbar_wait_susp(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t it = 0; it < (1u << 16); ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity, unsigned ns) {
    uint32_t ok = 0;
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (ns) __nanosleep(ns);
    }
    __trap();
}
#ifndef SP_WTC4_WAITMODE
#define SP_WTC4_WAITMODE 2   // bit 0: issuers park on lane 0 (measured 2.4x slower); bit 1: staging / drain warps use the hint
#endif
// one lane waits, the others park at the warp barrier
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) {
        if (SP_WTC4_WAITMODE & 2) mbar_wait_susp(bar, parity); else mbar_wait_poll(bar, parity, SP_WTC4_WAITMODE & 4 ? 200 : 32);
    }
    __syncwarp();
}
// the MMA-issuing warps: every lane polls (the loop stays warp-converged)
__device__ __forceinline__ void mbar_wait_issuer(uint32_t bar, uint32_t parity) {
    if (SP_WTC4_WAITMODE & 1) {
        if ((threadIdx.x & 31) == 0) mbar_wait_susp(bar, parity);
        __syncwarp();
    } else {
        mbar_wait_poll(bar, parity, SP_WTC4_WAITMODE & 8 ? 20 : 0);
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// v[8] -> two uint4 of bf16 terms, both rounded to nearest even with the packing conversion (one F2FP per channel pair
// and term): t1 = rn(v) differs from v by <= 2^-9 |v|, the residual is exact in fp32, t2 = rn(residual), so
// |v - t1 - t2| <= 2^-18 |v|.  nt == 1 (bf16 mode): the second term is zero.
__device__ __forceinline__ void split8_rn2(const float* v, uint4& o1, uint4& o2, int nt) {
    uint32_t p1[4], p2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        p1[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);                       // low 16 bits = even channel
        const float ra = v[2 * i] - __uint_as_float(p1[i] << 16);
        const float rb = v[2 * i + 1] - __uint_as_float(p1[i] & 0xffff0000u);
        p2[i] = pack_bf16x2(ra, rb);
    }
    o1 = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    o2 = (nt == 1) ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(p2[0], p2[1], p2[2], p2[3]);
}

// One work item = one tile column (n, 4 output rows, 32 input columns) of one slice pair (16 input x 16 output channels of the
// layer).  Items are numbered pair-major and dealt out round-robin (CTA b takes b, b + grid, ...: CTAs running side by side
// work on neighbouring tile columns, whose halo rows and dZ columns they share through L2 — contiguous ranges per CTA measured
// 1.15 -> 1.40 ms on the 16 -> 16 layer).  The accumulators are flushed to ws[pair][cta] at every pair change,
// wgrad_reduce_pairs_kernel folds the partials of a pair in CTA order.
struct ItemGeo {
    int n, oh0, u0;             // tile column
    int pair, cs, ci0, co0;     // slice pair, input slice index, first input / output channel
    int cie, coe;               // channels of the slices (<= 16)
};
// The staging loops walk their items in increasing order and keep the pair-derived fields in a PairWalk: a compare per item, the
// divisions only at a pair change (with a division by total_cols per item the 16 -> 16 layer measured 1.36 instead of 1.16 ms:
// the staging warps' address arithmetic is on the critical path of the pipeline although they wait for buffers most of the time).
struct PairWalk {
    int pair, lo, cs, ci0, co0, cie, coe;
};
__device__ __forceinline__ PairWalk pair_walk_init(int nsi, int Ci, int Co) {
    PairWalk w;
    w.pair = 0; w.lo = 0; w.cs = 0; w.ci0 = 0; w.co0 = 0;
    w.cie = Ci < 16 ? Ci : 16;
    w.coe = Co < 16 ? Co : 16;
    return w;
}
__device__ __forceinline__ void pair_walk_to(PairWalk& w, int item, int total_cols, int nsi, int Ci, int Co) {
    if (item < w.lo + total_cols) return;
    while (item >= w.lo + total_cols) { w.lo += total_cols; ++w.pair; }
    w.cs = w.pair % nsi;
    w.ci0 = w.cs * 16;
    w.co0 = (w.pair / nsi) * 16;
    w.cie = Ci - w.ci0 < 16 ? Ci - w.ci0 : 16;
    w.coe = Co - w.co0 < 16 ? Co - w.co0 : 16;
}
// MULTI == false: the layer is ONE pair (Ci, Co <= 16) — no pair arithmetic at all, the slice bounds are the kernel parameters
// (measured on the 16 -> 16 layer: 1.16 ms against 1.36 ms through the general form).
template <bool MULTI>
__device__ __forceinline__ ItemGeo item_geo(int item, PairWalk w, int total_cols, int tiles_w, int tiles_h, int nsi, int Ci, int Co) {
    ItemGeo c;
    int col = item;
    if (MULTI) {
        pair_walk_to(w, item, total_cols, nsi, Ci, Co);
        col = item - w.lo;
        c.pair = w.pair; c.cs = w.cs; c.ci0 = w.ci0; c.co0 = w.co0; c.cie = w.cie; c.coe = w.coe;
    } else {
        c.pair = 0; c.cs = 0; c.ci0 = 0; c.co0 = 0; c.cie = Ci; c.coe = Co;
    }
    const int tw = col % tiles_w;
    col /= tiles_w;
    c.oh0 = (col % tiles_h) * THW;
    c.n = col / tiles_h;
    c.u0 = tw * TU;
    return c;
}
__device__ __forceinline__ void drain_bar_sync() { asm volatile("bar.sync 1, 96;" ::: "memory"); }   // the three drain warps

template <bool MULTI>
__global__ void __launch_bounds__(NTHREADS4, 1)
wgrad3_tc4_kernel(SpConvDesc d, int nPerG, int G, int tiles_w, int tiles_h, int total_cols, int nsi, int items_total, int drain_every,
                  const float* __restrict__ X, const float* __restrict__ i_scale, const float* __restrict__ i_shift,
                  const float* __restrict__ dZ, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                  float* __restrict__ ws, long long* __restrict__ prof, int nt) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* a_reg = smem_raw;                                              // [buf][term][kw][half][row][u] x 16 B
    unsigned char* x_reg = smem_raw + A_REGION_B;                                 // [slot][row][term][half][u] x 16 B
    float* acc = reinterpret_cast<float*>(smem_raw + A_REGION_B + X_REGION_B);    // [(term_y, kw, co)][ACC_LD]
    float* coef = reinterpret_cast<float*>(smem_raw + A_REGION_B + X_REGION_B + ACC_B);   // [g][scale 16 | shift 16]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + A_REGION_B + X_REGION_B + ACC_B + COEF_B);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool pr = (prof != nullptr) && (blockIdx.x == 0);
    long long pw0 = 0, pw1 = 0, pwk = 0;

    for (int i = tid; i < ACC_ROWS * ACC_LD; i += NTHREADS4) acc[i] = 0.f;
    for (int i = tid; i < G * nsi * 32; i += NTHREADS4) {           // [g][input slice][scale 16 | shift 16]
        const int g = i / (nsi * 32), j = i & 31, c = ((i >> 5) % nsi) * 16 + (j & 15);
        float v = (j < 16) ? 1.f : 0.f;
        if (i_scale && c < d.Ci) v = (j < 16) ? i_scale[(int64_t)g * d.Ci + c] : i_shift[(int64_t)g * d.Ci + c];
        coef[i] = v;
    }
    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(smem_u32(&bars[b]), W_STG_G);                    // a_full[b]: one arrival per staging warp of group b
            mbar_init(smem_u32(&bars[NBUF + b]), NBLK);                // a_empty[b]: one commit per issuer
        }
        for (int b = 0; b < NBLK; ++b) {
            mbar_init(smem_u32(&bars[2 * NBUF + b]), 1);               // t_full[b]: issuer b
            mbar_init(smem_u32(&bars[2 * NBUF + NBLK + b]), 3);        // t_empty[b]: drain warps 0..2
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
    const long long t_start = pr ? clock64() : 0;                      // prof[10..13]: CTA 0 time line (issuer done, drain done, exit)

    const int first = (int)blockIdx.x, istep = (int)gridDim.x;         // this CTA's items: first + cl * istep
    const int ncols = (items_total - first + istep - 1) / istep;
    const int Do = d.Do;
    const int nsteps = ncols * Do;
    // last step of a slice pair inside this CTA: accumulators are drained and the partial is flushed
    // (no division per step in the issuing / draining loops: they walk (cl, od) incrementally, the pair test runs once per column)
    auto pair_end = [&](int cl, int od) {
        return od == Do - 1 && (cl == ncols - 1 || (first + (cl + 1) * istep) / total_cols != (first + cl * istep) / total_cols);
    };
    const int PPC = Do + 2;                                             // planes per column in the ring's sequence numbering

    if (warp >= W_EPI + W_MMA) {
        // =================================================================== staging: group grp takes the steps it = grp (mod NGRP)
        // Fixed roles inside a group, so a step costs pointer increments instead of index arithmetic: warps 0..3 stage the X'
        // planes (thread = (row parity, channel half, column), rows hy0, hy0 + 2, hy0 + 4), warps 4..6 the dZ tile (thread =
        // (channel half, dZ column j), four rows; 68 of 96 threads).  Lanes of a warp hold consecutive columns: 16-byte
        // shared-memory stores of a warp fall into consecutive banks.
        const int sw = warp - (W_EPI + W_MMA);
        const int grp = warp_uniform(sw / W_STG_G);
        const int wig = warp_uniform(sw % W_STG_G);                       // warp inside the group
        const bool isx = wig < W_STG_X;
        const int st = (wig - (isx ? 0 : W_STG_X)) * 32 + lane;           // thread inside its role
        const bool vec_i = (d.ldi % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
        const bool vec_o = (d.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(dZ) & 15) == 0);
        const int64_t xs_n = (int64_t)d.Di * d.Hi * d.Wi * d.ldi, zs_n = (int64_t)Do * d.Ho * d.Wo * d.ldo;
        // X' role
        const int wx = st & 31, xhalf = (st >> 5) & 1, hy0 = st >> 6;
        // dZ role
        const bool zact = !isx && st < 2 * ZW;
        const int zhalf = st >= ZW ? 1 : 0, zj = st - zhalf * ZW;
        int cl = grp / Do, od = grp - cl * Do, use = 0;
        PairWalk pw = pair_walk_init(nsi, d.Ci, d.Co);
        for (int it = grp; it < nsteps; it += NGRP, ++use) {
            const int buf = grp;                                         // NGRP == NBUF: group = buffer, use = it / NBUF
            while (od >= Do) { od -= Do; ++cl; }
            if (MULTI) pair_walk_to(pw, first + cl * istep, total_cols, nsi, d.Ci, d.Co);
            const ItemGeo cg = item_geo<MULTI>(first + cl * istep, pw, total_cols, tiles_w, tiles_h, nsi, d.Ci, d.Co);
            bool waited = false;
            long long c1 = pr ? clock64() : 0;
            if (isx) {
                // Planes staged by this step, in ring sequence numbers (column cl holds numbers cl*PPC + 0 .. Do+1, plane j of a
                // column is input depth j - pd).  Always plane od+2 of this column; the CTA's very first step also brings planes
                // 0 and 1; the last two steps of a column bring plane 0 / 1 of the NEXT column.  Safety of the ring of NBUF + 4
                // slots: a step's window is the three numbers from its base; the highest number a step writes (base + 4, a
                // next-column plane) replaces number base - NBUF, last read by the step NBUF before — whose MMAs this step has
                // waited for (a_empty of its buffer); the base advances by one per step (by three across a column change: Do + 2
                // numbers per column), so every other write replaces an older plane still.
                int np = 1;
                bool next1 = false;
                if (it == 0) np = 3;
                else if (od >= Do - 2 && cl + 1 < ncols) { np = 2; next1 = true; }
#pragma unroll 1
                for (int p = 0; p < np; ++p) {
                    const bool nx = (p == 1) && next1;
                    const ItemGeo c = nx ? item_geo<MULTI>(first + (cl + 1) * istep, pw, total_cols, tiles_w, tiles_h, nsi, d.Ci, d.Co) : cg;
                    const int pj = (p == 0) ? od + 2 : nx ? od - (Do - 2) : p - 1;          // plane of its column
                    const int seq = (nx ? cl + 1 : cl) * PPC + pj;
                    const int gd = pj - d.pd, gw = c.u0 + wx, gh0 = c.oh0 - d.ph + hy0;
                    const int ch = xhalf * 8;
                    const bool okp = gd >= 0 && gd < d.Di && gw < d.Wi && ch < c.cie;
                    const float* pp = X + (int64_t)c.n * xs_n + (((int64_t)gd * d.Hi + gh0) * d.Wi + gw) * d.ldi + c.ci0 + ch;
                    const int64_t rstep = (int64_t)2 * d.Wi * d.ldi;
                    float4 ra[3], rb[3];
                    bool ok[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int gh = gh0 + 2 * i;
                        ok[i] = okp && gh >= 0 && gh < d.Hi;
                        ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        rb[i] = ra[i];
                        if (ok[i]) {
                            const float* q = pp + i * rstep;
                            if (vec_i && ch + 8 <= c.cie) {
                                ra[i] = *reinterpret_cast<const float4*>(q);
                                rb[i] = *reinterpret_cast<const float4*>(q + 4);
                            } else {
                                float e[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) e[j] = (ch + j < c.cie) ? q[j] : 0.f;
                                ra[i] = make_float4(e[0], e[1], e[2], e[3]);
                                rb[i] = make_float4(e[4], e[5], e[6], e[7]);
                            }
                        }
                    }
                    // convert BEFORE the buffer is waited for: after the wait only the stores remain (the time from "MMAs of step
                    // it - NBUF done" to "step it staged" is what the issuers see as staging latency)
                    const float* cf = coef + ((c.n / nPerG) * nsi + c.cs) * 32 + ch;
                    const float4 s0 = *reinterpret_cast<const float4*>(cf), s1 = *reinterpret_cast<const float4*>(cf + 4);
                    const float4 h0 = *reinterpret_cast<const float4*>(cf + 16), h1 = *reinterpret_cast<const float4*>(cf + 20);
                    uint4 o1[3], o2[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
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
                    unsigned char* dp = x_reg + (seq % NSLOT) * X_PLANE_B + (hy0 * 4 + xhalf) * RS + wx * 16;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        *reinterpret_cast<uint4*>(dp + i * 8 * RS) = o1[i];
                        *reinterpret_cast<uint4*>(dp + i * 8 * RS + 2 * RS) = o2[i];
                    }
                }
            } else {
                const float* zn = dZ + (int64_t)cg.n * zs_n + cg.co0;
                const int gz = cg.n / nPerG;
                const int gw = cg.u0 + d.pw - 2 + zj, ch = zhalf * 8;
                const bool okc = zact && gw >= 0 && gw < d.Wo && ch < cg.coe;
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
                        const float* q = pp + i * rstep;
                        if (vec_o && ch + 8 <= cg.coe) {
                            ra[i] = *reinterpret_cast<const float4*>(q);
                            rb[i] = *reinterpret_cast<const float4*>(q + 4);
                        } else {
                            float e[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) e[j] = (ch + j < cg.coe) ? q[j] : 0.f;
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
                            if (ch + j < cg.coe) v[j] = fmaf(v[j], o_scale[(int64_t)gz * d.Co + cg.co0 + ch + j], o_shift[(int64_t)gz * d.Co + cg.co0 + ch + j]);
                    }
                    split8_rn2(v, o1[i], o2[i], nt);
                }
                if (pr) c1 = clock64();
                mbar_wait_warp(a_empty + 8 * buf, (use & 1) ^ 1);
                if (pr) { const long long c2 = clock64(); pw0 += c2 - c1; c1 = c2; }
                if (zact) {
                    unsigned char* dp = a_reg + buf * A_BUF_B + zhalf * PS + (zj - 2) * 16;      // + kw * (2 PS + 16) + row * TU * 16
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int k = zj - 2 + kw;               // A_kw[u] = dZ[u + pw - kw]
                        if (k >= 0 && k < TU) {
#pragma unroll
                            for (int i = 0; i < THW; ++i) {
                                unsigned char* q = dp + kw * (2 * PS + 16) + i * (TU * 16);
                                *reinterpret_cast<uint4*>(q) = o1[i];
                                *reinterpret_cast<uint4*>(q + 6 * PS) = o2[i];
                            }
                        }
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * buf);
            if (pr) pwk += clock64() - c1;
            od += NGRP;
        }
        if (pr && grp == 0 && wig == 0 && lane == 0) { prof[4] = pw0; prof[5] = pwk; }
        if (pr && grp == 0 && wig == W_STG_X && lane == 0) { prof[8] = pw0; prof[9] = pwk; }
    } else if (warp >= W_EPI) {
        // =================================================================== MMA issue: warp 4 + kd owns accumulator block kd
        // (the whole warp runs the loop with warp-uniform values, one elected lane issues: see umma_bf16_elect)
        {
            const int kd = warp_uniform(warp - W_EPI);
            const uint32_t a_base = smem_u32(a_reg), x_base = smem_u32(x_reg);
            const uint32_t dcol = tmem_base + (uint32_t)(kd * BCOLS);
            constexpr uint32_t IDESC = idesc_mn(128, BCOLS);
            bool fresh = true;
            int drains = 0, sp = 0;                                          // sp: steps since the last pair change
            int cl = 0, od = 0, buf = 0, use = 0, ring = kd;                // ring = (cl * PPC + od + kd) % NSLOT
            for (int it = 0; it < nsteps; ++it) {
                long long c0 = pr ? clock64() : 0;
                mbar_wait_issuer(a_full + 8 * buf, use & 1);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                if (fresh && drains > 0) mbar_wait_issuer(t_empty + 8 * kd, (drains - 1) & 1);
                long long c2 = pr ? clock64() : 0;
                pw1 += c2 - c1;
                tc_fence_after();
                const int slot = ring;                                // plane od - pd + kd of this column
                const uint64_t da0 = umma_desc(a_base + (uint32_t)(buf * A_BUF_B), 128, PS);
                const uint64_t db0 = umma_desc(x_base + (uint32_t)(slot * X_PLANE_B), 128, RS);
#pragma unroll
                for (int r = 0; r < THW; ++r) {
#pragma unroll
                    for (int ks = 0; ks < TU / 16; ++ks) {
                        const uint64_t da = da0 + (uint64_t)((r * TU * 16 + ks * 256) >> 4);
                        const uint64_t db = db0 + (uint64_t)((r * 4 * RS + ks * 256) >> 4);
                        umma_bf16_elect(dcol, da, db, IDESC, fresh ? 0u : 1u);
                        fresh = false;
                    }
                }
                umma_commit_elect(a_empty + 8 * buf);
                const bool pend = pair_end(cl, od);
                if (sp + 1 == drain_every || pend) {
                    umma_commit_elect(t_full + 8 * kd);
                    fresh = true;
                    ++drains;
                    sp = 0;
                } else {
                    ++sp;
                }
                if (pr) pwk += clock64() - c2;
                if (++buf == NBUF) { buf = 0; ++use; }
                int adv = 1;                                           // ring numbers to the next step's window: 3 across a column change
                if (++od == Do) { od = 0; ++cl; adv = 3; }
                ring += adv;
                if (ring >= NSLOT) ring -= NSLOT;
            }
            if (pr && kd == 0 && lane == 0) { prof[0] = pw0; prof[1] = pw1; prof[2] = pwk; prof[3] = nsteps; prof[10] = clock64() - t_start; }
        }
    } else if (warp < 3) {
        // =================================================================== drain: warp q holds accumulator rows 32q .. 32q+31
        float* arow = acc + (size_t)(warp * 32 + lane) * ACC_LD;
        int dr = 0;
        for (int cl = 0; cl < ncols;) {
            // run of this CTA's columns that belong to one slice pair: (run * Do) steps, a drain every drain_every steps and at its end
            const int item0 = first + cl * istep;
            const int pair = item0 / total_cols;
            int run = ((pair + 1) * total_cols - item0 + istep - 1) / istep;
            if (run > ncols - cl) run = ncols - cl;
            const int ndr = (run * Do + drain_every - 1) / drain_every;
            for (int k = 0; k < ndr; ++k, ++dr) {
#pragma unroll 1
                for (int kd = 0; kd < NBLK; ++kd) {
                    long long c0 = pr ? clock64() : 0;
                    mbar_wait_warp(t_full + 8 * kd, dr & 1);
                    long long c1 = pr ? clock64() : 0;
                    pw0 += c1 - c0;
                    tc_fence_after();
#pragma unroll 1
                    for (int kh = 0; kh < 3; ++kh) {
                        float v[32];                               // [x term 1: ci 0..15 | x term 2: ci 0..15]
                        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kd * BCOLS + kh * 32), v);
                        if (kh == 2) {                             // the block is in registers / shared memory: it may be overwritten
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
            // flush: fold the two y terms (small first) and write this CTA's partial of the pair as [co 16][ci 16][27]
            float* wsp = ws + ((int64_t)pair * istep + first) * WS_SLOT;
            drain_bar_sync();                                      // every drain warp has added its rows
            for (int i = warp * 32 + lane; i < WS_SLOT; i += 96) {
                const int tap = i % 27, ci = (i / 27) & 15, co = i / (27 * 16);
                const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                const int col = kd * 48 + kh * 16 + ci;
                const int r0 = (kw * 2 + (co >> 3)) * 8 + (co & 7);          // term 0; term 1 lies 48 rows below
                wsp[i] = acc[(r0 + 48) * ACC_LD + col] + acc[r0 * ACC_LD + col];
            }
            drain_bar_sync();                                      // rows are read across warps above
            for (int j = 0; j < ACC_LD; ++j) arow[j] = 0.f;        // own row only: the next drain of this thread follows in order
            cl += run;
        }
        if (pr && tid == 0) { prof[6] = pw0; prof[7] = pwk; prof[11] = clock64() - t_start; }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
    if (pr && tid == 0) prof[12] = clock64() - t_start;
}

// dw[co][ci][tap] (torch layout, the whole layer) = beta * dw + the per-CTA partials of every slice pair, folded in CTA order
__global__ void wgrad_reduce_pairs_kernel(const float* __restrict__ ws, int T, int grid, int nsi, int npairs, int Ci, int Co,
                                          float* __restrict__ dw, float beta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs * WS_SLOT) return;
    const int pair = i / WS_SLOT, e = i - pair * WS_SLOT;
    const int tap = e % 27, ci = (e / 27) & 15, co = e / (27 * 16);
    const int gci = (pair % nsi) * 16 + ci, gco = (pair / nsi) * 16 + co;
    if (gci >= Ci || gco >= Co) return;
    // the CTAs that hold items of the pair form a cyclic range: all of them when T >= grid, else T CTAs from (pair * T) % grid
    const int cnt = T >= grid ? grid : T;
    int b = T >= grid ? 0 : (int)(((int64_t)pair * T) % grid);
    const float* base = ws + (int64_t)pair * grid * WS_SLOT + e;
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < cnt; ++k) {                    // fixed order; the loads of a batch of iterations are independent
        s += base[(int64_t)b * WS_SLOT];
        if (++b == grid) b = 0;
    }
    float* o = dw + ((int64_t)gco * Ci + gci) * 27 + tap;
    *o = (beta != 0.f ? beta * *o : 0.f) + s;
}

struct Wtc4Plan {
    int tiles_w, tiles_h, grid, nsi, nso, npairs;
    int64_t total;      // tile columns (n, 4 output rows, 32 input columns) of ONE slice pair; each is walked along the depth axis
    int64_t items;      // npairs * total
};
static inline Wtc4Plan plan(const SpConvDesc* d) {
    Wtc4Plan p;
    p.tiles_w = (d->Wi + TU - 1) / TU;
    p.tiles_h = (d->Ho + THW - 1) / THW;
    p.total = (int64_t)p.tiles_w * p.tiles_h * d->N;
    p.nsi = (d->Ci + 15) / 16;
    p.nso = (d->Co + 15) / 16;
    p.npairs = p.nsi * p.nso;
    p.items = p.total * p.npairs;
    p.grid = sp_num_sms();
    const int cap = sp_wtc4_grid_cap_ref();   // tests: several columns per CTA on a small geometry
    if (cap > 0 && p.grid > cap) p.grid = cap;
    if (p.grid > p.items) p.grid = (int)p.items;
    if (p.grid < 1) p.grid = 1;
    return p;
}

}  // namespace sp_wtc4

static inline bool sp_tc4_wgrad_disabled() { return sp_wtc4_generation_ref() != 2; }

// 3x3x3 stride-1 layers with 2..96 input and 9..64 output channels: ONE launch walks the tile columns of every 16 x 16 channel
// slice pair (Cae3D.py:44,208,211 and Unet3D.py:22: one pair; Cae3D.py:52,55,186-200: 2..4 pairs; Unet3D.py:19,22: up to 16).
// G = statistics groups of the launch.
static inline bool sp_tc4_wgrad_supported(const SpConvDesc* d, int G) {
    if (sp_tc4_wgrad_disabled() || G < 1 || G > sp_wtc4::MAXG) return false;
    if (d->k != 3 || d->s != 1 || sp_tc_terms() == 0 || sp_tc_wgrad_disabled()) return false;
    // 2..8 input channels (Unet3D.py:19 block1: 2 -> 16; Enc3DCtp: 3 channels): the upper channel half of X' is staged as zeros —
    // the MMA work is wasted but the kernel is bound by staging, not by the tensor pipe: 1.22 -> 0.70 ms against the thin-input
    // kernel on the U-Net's first layer; a single input channel (Cae3D.py:41) stays on sp_conv_thin.cuh (0.83 vs 0.9 ms)
    if (d->Ci < 2 || d->Ci > 16 * sp_wtc4::MAXSI || d->Co <= 8 || d->Co > 64) return false;
    if (d->Ci > 16 || d->Co > 16) {          // slices: 16-byte aligned channel slices; no slice of fewer than ... channels is excluded
        if (d->ldi % 4 != 0 || d->ldo % 4 != 0) return false;
        static int maxpairs = -1;            // SP_WTC4_SLICE_PAIRS: most slice pairs taken (every pair re-stages both tiles)
        if (maxpairs < 0) {
            const char* e = getenv("SP_WTC4_SLICE_PAIRS");
            maxpairs = e ? atoi(e) : 16;
        }
        if (((d->Ci + 15) / 16) * ((d->Co + 15) / 16) > maxpairs) return false;
    }
    if (d->pd > 2 || d->ph > 2 || d->pw > 2) return false;
    const sp_wtc4::Wtc4Plan p = sp_wtc4::plan(d);
    return p.total >= 16 && p.items < (1LL << 31) / (d->Do + 2) && d->Wo >= 24 && d->Do >= 4;
}

static inline size_t sp_tc4_wgrad_workspace_bytes(const SpConvDesc* d) {
    const sp_wtc4::Wtc4Plan p = sp_wtc4::plan(d);
    return (size_t)p.grid * p.npairs * sp_wtc4::WS_SLOT * sizeof(float);
}

// pointers of sliced layers must be 16-byte aligned (float4 loads of a channel slice)
static inline bool sp_tc4_wgrad_aligned(const SpConvDesc* d, const void* iside, const void* oside) {
    if (d->Ci <= 16 && d->Co <= 16) return true;
    return ((reinterpret_cast<uintptr_t>(iside) | reinterpret_cast<uintptr_t>(oside)) & 15) == 0;
}

static inline int sp_tc4_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                      const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                      cudaStream_t st, long long* prof = nullptr, int drain_every = 6) {   // 48 accumulations per drain: rel-L2 4.6e-7 (8 steps: 6.7e-7 and 4 % faster; 2 steps: 3.0e-7)
    using namespace sp_wtc4;
    const Wtc4Plan p = plan(d);
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(wgrad3_tc4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM4));
        SP_CUDA(cudaFuncSetAttribute(wgrad3_tc4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM4));
        attr = true;
    }
    const int G = d->N / nPerG;
    if (p.npairs == 1)
        wgrad3_tc4_kernel<false><<<p.grid, NTHREADS4, SMEM4, st>>>(*d, nPerG, G, p.tiles_w, p.tiles_h, (int)p.total, p.nsi, (int)p.items, drain_every,
                                                                   iside, i_scale, i_shift, oside, o_scale, o_shift, ws, prof, sp_tc_terms() == 1 ? 1 : 2);
    else
        wgrad3_tc4_kernel<true><<<p.grid, NTHREADS4, SMEM4, st>>>(*d, nPerG, G, p.tiles_w, p.tiles_h, (int)p.total, p.nsi, (int)p.items, drain_every,
                                                                  iside, i_scale, i_shift, oside, o_scale, o_shift, ws, prof, sp_tc_terms() == 1 ? 1 : 2);
    SP_LAUNCH_OK("wgrad3_tc4_kernel");
    const int n = p.npairs * WS_SLOT;
    wgrad_reduce_pairs_kernel<<<(n + 255) / 256, 256, 0, st>>>(ws, (int)p.total, p.grid, p.nsi, p.npairs, d->Ci, d->Co, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_pairs_kernel");
    return 0;
}
