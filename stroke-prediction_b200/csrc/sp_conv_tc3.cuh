// sp_conv_tc3.cuh — third-generation tcgen05 / TMEM correlation for the 3x3x3 stride-1 layers (Cae3D.py:44,52,55,186-211;
// Unet3D.py:19,22) and, through flipped taps, their dgrads.  Same split-bf16 arithmetic as sp_conv_tc2.cuh (three bf16 terms
// per fp32 operand, products of order <= 2, leading products and corrections in separate accumulators), restructured around the
// two things ncu showed binding generation 2 (tensor pipe 12-23 % active: staging and MMA issue, not the pipe):
//
//  * kw-stacked N.  The three taps along w share ONE A tile: D[voxel, (kw, co)] = sum_ci X[voxel][ci] * W[kd,kh,kw][co][ci], so a
//    (kd, kh) pair is ONE MMA per activation term with N = 3 x Co x terms (144 / 96 / 48 columns for 16 channels) instead of
//    three with N = 48 / 32 / 16: 27 MMAs per 128-row tile instead of 81, each long enough (72 / 48 / 24 pipe cycles) to keep the
//    tensor pipe fed.  The epilogue adds the three kw partial sums from NEIGHBOURING rows (out[w] = D0[w] + D1[w+1] + D2[w+2]):
//    rows are (h, w'') with w'' fastest, so that is two warp shuffles per column; a 16-wide row yields 14 outputs.
//    Accuracy is unchanged: every leading accumulator still takes 9 accumulations (kd, kh), the epilogue adds 3 of them in RN.
//  * rolling depth window.  A CTA owns a (8 rows x 14 columns) output patch and walks along depth: every input plane (10 x 16
//    halo voxels) is staged ONCE into a ring of 9 (bf16 mode: 16) shared-memory planes and used by the three output planes around it (generation 2
//    staged 4 halo planes per 2 output planes: 2.0x the loads, splits and stores per output).  The pipeline never drains: ring,
//    accumulator and barrier phases run on global counters across the CTA's work items.
//
//  * TMA staging.  One thread streams the raw fp32 halo planes with cp.async.bulk.tensor (5-D tensor map of the NDHWC source,
//    box [16 c][16 w][10 h]): zero padding, ragged edge tiles and channels >= Ci arrive as zeros, the staging warps only convert
//    (BatchNorm, split) from shared memory to the UMMA layout — no address arithmetic, bounds checks or global-load latency.
//
// Roles (one persistent CTA per SM, 22 warps; the last one is the TMA producer): warps 0-7 epilogue (two groups of four alternate output planes; TMEM lane
// quarter = warp % 4), warps 8-17 staging (one (voxel, 8-channel chunk) item per thread and plane, the next plane's global loads
// in flight while the current one is split and stored), warps 18-20 one MMA-issuing thread each (output plane q -> issuer
// q % NACC, accumulator q % NACC: MMAs of one thread retire one after the other, different issuers overlap).
// NS = 1 is the bf16 mode: operands rounded to ONE bf16 term (RN), one MMA per (kd, kh), fp32 accumulation.
#pragma once
#include <cuda.h>
#include "sp_conv_tc2.cuh"

namespace sp_tc3 {

using namespace sp_tc;
using sp_tc2::mbar_arrive;
using sp_tc2::split8_trunc3;
using sp_tc2::tmem_ld_n;

constexpr int TW = 16, TWV = 14, TH = 8, IH = TH + 2;   // MMA rows = 16 w'' x 8 h; 14 x 8 outputs per plane
constexpr int PSLOTS = IH * TW;                          // 160 sixteen-byte slots per (term, chunk) plane
constexpr int CHS = PSLOTS + 4;                          // chunk-plane stride: odd multiple of 64 B (bank spread of the 2 K chunks)
#ifndef SP_TC3_NEPI_G
#define SP_TC3_NEPI_G 2
#endif
constexpr int NEPI_G = SP_TC3_NEPI_G, NEPI_W = 4 * NEPI_G;           // epilogue groups / warps (three groups measured slower: 25 warps crowd the
                                                          // schedulers the three MMA-issuing threads need: issue time 3.7 k -> 5.2 k cycles per plane)
constexpr int NSTG_W = 10, NSTG = NSTG_W * 32;           // staging warps / threads: 320 = one item per thread and plane
constexpr int NISS_W = 3;                                // MMA issuer warps (NACC of them active)
constexpr int NWARPS3 = NEPI_W + NSTG_W + NISS_W + 1;    // 22: + the TMA producer warp
constexpr int NTHREADS3 = NWARPS3 * 32;                  // 704
constexpr int RAW_B = PSLOTS * 16 * 4;                   // one raw fp32 halo plane as TMA lands it: [10 h][16 w][16 c] = 10240 B
static_assert(PSLOTS * 2 == NSTG, "one (slot, chunk) item per staging thread");

template <int COP, int NS>
struct Tc3 {
    static constexpr int TS = (3 * COP + 15) / 16 * 16;          // N rows / TMEM columns per weight term: 3 kw x COP, padded to 16
    static constexpr int NTOT = NS * TS;                         // rows of the weight image per ((kd, kh), chunk) = N of the a1 MMA
    static constexpr int BROWS = NTOT + 4;                       // padded chunk stride of the image in shared memory
    static constexpr int ACOLS = NTOT;                           // TMEM columns per output plane: [main | cA | cB]
    static constexpr int NACC = (512 / ACOLS) < 3 ? (512 / ACOLS) : 3;
    static constexpr int SLOT_U4 = NS * 2 * CHS;                 // uint4 per ring slot: [term][chunk][CHS]
    // staged input planes in flight: NACC issuers hold NACC + 2 planes; the rest is the stager's lead over them (measured with
    // 6: issuers waited for planes 25 % of the time while the stager waited for slots)
    static constexpr int RING = NS == 3 ? (COP <= 16 ? 9 : 7) : 16;
    // planes of global loads a staging thread keeps in flight (registers): bf16 mode is HBM-bound at ~740 cycles per plane,
    // below the ~1.5 k cycle load latency under load
    static constexpr int PF = NS == 3 ? 2 : 4;
    static constexpr int WIMG = 9 * 2 * NTOT;                    // global image (uint4)
    static constexpr int WIMGS = 9 * 2 * BROWS;                  // shared-memory image (uint4)
    // t_full barriers: plane q is drained by epilogue group q % NEPI_G from accumulator q % NACC.  A group must see EVERY phase
    // of a barrier it waits on (a parity wait cannot tell phase p from p + 2), so there is one barrier per (accumulator, group)
    // combination = q % lcm(NACC, NEPI_G), each completing once per NTF planes, always awaited by the same group.
    static constexpr int NTF = (NACC % NEPI_G == 0) ? NACC : NACC * NEPI_G;
    // raw fp32 planes in flight between the TMA producer and the staging warps (bf16 mode is HBM-bound: deeper)
    static constexpr int KR = NS == 3 ? 4 : 8;
    static constexpr int NBARS = 2 * RING + NTF + 3 + 2 * KR;
    static constexpr size_t SMEM = (size_t)KR * RAW_B + ((size_t)RING * SLOT_U4 + WIMGS) * 16 + NBARS * 8 + 64;   // + barriers + TMEM slot
    static_assert(NTOT % 16 == 0 && NTOT <= 256, "UMMA M = 128 needs N % 16 == 0, N <= 256");
    static_assert(NACC >= 2, "two accumulators in flight at least");
};
static_assert(Tc3<16, 3>::SMEM <= 227 * 1024 && Tc3<24, 3>::SMEM <= 227 * 1024 && Tc3<24, 1>::SMEM <= 227 * 1024, "tc3: shared memory");

// ---- weight image -----------------------------------------------------------------------------------------------------
// img[((kd*3+kh) * 2 + chunk) * NTOT + term * TS + kw * COP + n] = 8 bf16 {term of Wsrc(n0 + n, k0 + chunk*8 + j, tap)}, zero rows
// between 3*COP and TS.  transposed: GEMM N = conv-ci, K = conv-co, taps flipped (sp_corrT as a correlation).
template <int NS>
__global__ void pack_wimg3_kernel(const float* __restrict__ w, int Co, int Ci, int transposed, int COP, int TS, int k0, int n0,
                                  uint4* __restrict__ img) {
    const int NTOT = NS * TS;
    const int total = 9 * 2 * TS;
    const int Nn = transposed ? Ci : Co, Kk = transposed ? Co : Ci;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int row = i % TS;                       // kw * COP + n, or padding
        const int chunk = (i / TS) % 2;
        const int t9 = i / (2 * TS);
        const int kw = row / COP, nl = row % COP;
        const int n = n0 + nl;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + chunk * 8 + j;
            float x = 0.f;
            if (kw < 3 && n < Nn && k < Kk) {
                const int tap = t9 * 3 + kw;
                x = transposed ? w[((int64_t)k * Ci + n) * 27 + (26 - tap)] : w[((int64_t)n * Ci + k) * 27 + tap];
            }
            v[j] = x;
        }
        uint4 o[3];
        if (NS == 3) {
            split8<3>(v, o);
        } else {
            o[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) img[(t9 * 2 + chunk) * NTOT + s * TS + row] = o[s];
    }
}

template <int N>
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t* r) {
    static_assert(N % 8 == 0, "tmem_ld_issue: multiples of 8 columns");
    constexpr int N16 = N / 16 * 16;
#pragma unroll
    for (int c = 0; c < N16; c += 16)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[c]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]), "=r"(r[c + 6]), "=r"(r[c + 7]),
                       "=r"(r[c + 8]), "=r"(r[c + 9]), "=r"(r[c + 10]), "=r"(r[c + 11]), "=r"(r[c + 12]), "=r"(r[c + 13]), "=r"(r[c + 14]), "=r"(r[c + 15])
                     : "r"(taddr + c) : "memory");
#pragma unroll
    for (int c = N16; c < N; c += 8)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[c]), "=r"(r[c + 1]), "=r"(r[c + 2]), "=r"(r[c + 3]), "=r"(r[c + 4]), "=r"(r[c + 5]), "=r"(r[c + 6]), "=r"(r[c + 7])
                     : "r"(taddr + c) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Item3 { int n, oh0, ow0, od_lo, L; };

// Diagnostics (tools/tc_probe): with SP_TC3_DEBUG a wait that times out records who waited for what and lets the kernel run to
// its end with garbage instead of trapping, so the host can read the record.  [0] = code of the first timeout, [1] = its counter.
__device__ unsigned long long sp_tc3_dbg[4];

// Warp-collective wait: ONE lane polls the barrier (with a short sleep between polls), the others park at the warp barrier.
// Measured with every thread spinning: a third of all issued instructions of the kernel were SYNCS / BRA of 20 waiting warps,
// taken from the issue slots of the epilogue warps that share their schedulers.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int dbg, unsigned code, unsigned counter);

__device__ __forceinline__ void mbar_wait3(uint32_t bar, uint32_t parity, int dbg, unsigned code, unsigned counter) {
    if (!dbg) {
        uint32_t ok = 0;
        for (uint32_t it = 0; it < (1u << 24); ++it) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            if (ok) return;
            __nanosleep(40);
        }
        __trap();
    }
    if (sp_tc3_dbg[0] != 0ull) return;                    // somebody timed out already: drain
    uint32_t ok = 0;
    for (uint32_t it = 0; it < (1u << 20); ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if ((it & 1023) == 1023 && sp_tc3_dbg[0] != 0ull) return;
    }
    if (atomicCAS(&sp_tc3_dbg[0], 0ull, (unsigned long long)code | ((unsigned long long)blockIdx.x << 32)) == 0ull)
        sp_tc3_dbg[1] = counter | ((unsigned long long)parity << 32);
}

__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int dbg, unsigned code, unsigned counter) {
    if ((threadIdx.x & 31) == 0) mbar_wait3(bar, parity, dbg, code, counter);
    __syncwarp();
}

// accum != 0: add the raw sums already in dst (later input-channel passes); fin == 0: store raw sums (no bias, no activation).
// sstride: floats between the scale / shift rows of two statistics groups (the layer's full channel count when src is a slice).
template <int COP, int NS>
__global__ void __launch_bounds__(NTHREADS3, 1)
corr3_tc3_kernel(SpConvDesc d, int nPerG, int tiles_w, int tiles_h, int nseg, int seg_len, int total_items,
                 const float* __restrict__ src, const uint4* __restrict__ wimg, const float* __restrict__ bias,
                 const float* __restrict__ scale, const float* __restrict__ shift, int sstride, int accum, int fin,
                 float* __restrict__ dst, long long* __restrict__ prof, int dbg, const __grid_constant__ CUtensorMap tmap, int use_tma) {
    using T = Tc3<COP, NS>;
    constexpr int TS = T::TS, NTOT = T::NTOT, BROWS = T::BROWS, ACOLS = T::ACOLS, NACC = T::NACC, SLOT_U4 = T::SLOT_U4;
    constexpr int WIMG = T::WIMG, WIMGS = T::WIMGS, RING = T::RING, PF = T::PF, NTF = T::NTF, KR = T::KR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* raw = smem_raw;                                          // [KR][10 h][16 w][16 c] fp32, written by TMA
    uint4* As = reinterpret_cast<uint4*>(smem_raw + (size_t)KR * RAW_B);     // [RING][NS][2][CHS]
    uint4* Bs = As + (size_t)RING * SLOT_U4;                                 // weight image [9][2][BROWS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + WIMGS);                 // a_full[RING] a_empty[RING] t_full[NTF] t_empty[3]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + T::NBARS);
    const bool pr = (prof != nullptr) && (blockIdx.x == 0);
    long long pw0 = 0, pw1 = 0, pwk = 0;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < WIMG; i += NTHREADS3) Bs[(i / NTOT) * BROWS + (i % NTOT)] = wimg[i];
    if (tid == 0) {
        for (int s = 0; s < RING; ++s) {
            mbar_init(smem_u32(&bars[s]), NSTG_W);               // a_full: one arrival per staging warp
            mbar_init(smem_u32(&bars[RING + s]), 3);             // a_empty: one tcgen05.commit per output plane that read the slot
        }
        for (int a = 0; a < NTF; ++a) mbar_init(smem_u32(&bars[2 * RING + a]), 1);         // t_full: the issuer of the plane
        for (int a = 0; a < 3; ++a) mbar_init(smem_u32(&bars[2 * RING + NTF + a]), 4);     // t_empty: the 4 warps of the draining group
        for (int k = 0; k < KR; ++k) {
            mbar_init(smem_u32(&bars[2 * RING + NTF + 3 + k]), 1);                          // raw_full: the producer's expect_tx arrival
            mbar_init(smem_u32(&bars[2 * RING + NTF + 3 + KR + k]), NSTG_W);                // raw_empty: one arrival per staging warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[RING]);
    const uint32_t t_full = smem_u32(&bars[2 * RING]), t_empty = smem_u32(&bars[2 * RING + NTF]);
    const uint32_t raw_full = smem_u32(&bars[2 * RING + NTF + 3]), raw_empty = smem_u32(&bars[2 * RING + NTF + 3 + KR]);

    auto item_of = [&](int item) {
        Item3 it;
        int t = item;
        const int sg = t % nseg; t /= nseg;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th_ = t % tiles_h;
        it.n = t / tiles_h;
        it.ow0 = tw * TWV; it.oh0 = th_ * TH;
        it.od_lo = sg * seg_len;
        it.L = (d.Do - it.od_lo < seg_len) ? d.Do - it.od_lo : seg_len;
        return it;
    };

    if (warp == NWARPS3 - 1) {
        // =================================================================== TMA producer: one thread streams the raw halo planes
        // One cp.async.bulk.tensor per input plane: box [16 c][16 w][10 h] of the fp32 NDHWC source at (signed) coordinates
        // (0, ow0 - pw, oh0 - ph, id, n); everything outside the tensor — the zero padding of the convolution, ragged edge tiles,
        // channels >= Ci — arrives as zeros.  No address arithmetic, bounds checks or load latency in the staging warps.
        if (lane == 0 && use_tma) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
            uint32_t gp = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                const Item3 it = item_of(item);
                const int c_w = it.ow0 - d.pw, c_h = it.oh0 - d.ph, c_d0 = it.od_lo - d.pd;
                for (int ip = 0; ip < it.L + 2; ++ip, ++gp) {
                    const uint32_t k = gp % KR, use = gp / KR;
                    mbar_wait3(raw_empty + 8 * k, (use & 1) ^ 1, dbg, 0x500u + k, gp);      // the staging warps have read this slot
                    const uint32_t bar = raw_full + 8 * k, dsts = smem_u32(raw + (size_t)k * RAW_B);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)RAW_B) : "memory");
                    asm volatile(
                        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                        :: "r"(dsts), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(c_w), "r"(c_h), "r"(c_d0 + ip), "r"(it.n), "r"(bar)
                        : "memory");
                }
            }
        }
    } else if (warp >= NEPI_W && warp < NEPI_W + NSTG_W) {
        // =================================================================== staging warps
        const int st = tid - NEPI_W * 32;                        // 0..319
        const int chunk = st & 1, slot = st >> 1;                // this thread's item of every plane
        const int hy = slot / TW, wx = slot % TW;
        const int c = chunk * 8;
        const bool vec = (d.ldi % 4 == 0);
        uint32_t gin = 0;                                        // global input-plane counter of this CTA
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item3 it = item_of(item);
            const int g = it.n / nPerG;
            float bsc[8], bsh[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool okc = scale && (c + j < d.Ci);
                bsc[j] = okc ? scale[(int64_t)g * sstride + c + j] : 1.f;
                bsh[j] = okc ? shift[(int64_t)g * sstride + c + j] : 0.f;
            }
            const int gh = it.oh0 - d.ph + hy, gw = it.ow0 - d.pw + wx;
            const bool in_hw = gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi && c < d.Ci;
            const float* colp = src + (((int64_t)it.n * d.Di * d.Hi + gh) * d.Wi + gw) * d.ldi + c;   // + gd * Hi*Wi*ldi
            const int64_t dstride = (int64_t)d.Hi * d.Wi * d.ldi;
            const int id0 = it.od_lo - d.pd;
            const int nplanes = it.L + 2;
            auto load_plane = [&](int ip, float4& a, float4& b) -> bool {
                const int gd = id0 + ip;
                a = make_float4(0.f, 0.f, 0.f, 0.f);
                b = a;
                if (!(in_hw && gd >= 0 && gd < d.Di)) return false;
                const float* p = colp + (int64_t)gd * dstride;
                if (vec && c + 8 <= d.Ci) {
                    a = *reinterpret_cast<const float4*>(p);
                    b = *reinterpret_cast<const float4*>(p + 4);
                } else {
                    float e[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) e[j] = (c + j < d.Ci) ? p[j] : 0.f;
                    a = make_float4(e[0], e[1], e[2], e[3]);
                    b = make_float4(e[4], e[5], e[6], e[7]);
                }
                return true;
            };
            if (use_tma) {
                // raw plane from the TMA ring -> (BN) -> split -> UMMA layout.  Whether a voxel is real (BatchNorm applies) or padding
                // (stays exactly zero after BatchNorm, Cae3D.py:40-41) is pure coordinate arithmetic.
#pragma unroll 1
                for (int ip = 0; ip < nplanes; ++ip, ++gin) {
                    const uint32_t k = gin % KR, ruse = gin / KR;
                    mbar_wait_warp(raw_full + 8 * k, ruse & 1, dbg, 0x600u + k, gin);
                    const float4* rp = reinterpret_cast<const float4*>(raw + (size_t)k * RAW_B + (size_t)(slot * 16 + c) * 4);
                    const float4 ra = rp[0], rb = rp[1];
                    // generic-proxy reads before the async-proxy (TMA) write that refills this slot: a cross-proxy WAR needs the proxy
                    // fence as well as the barrier (without it ~4 % of the launches had a few stale voxels)
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(raw_empty + 8 * k);           // the raw slot may be refilled
                    const int gd = id0 + ip;
                    float v[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                    if (scale && in_hw && gd >= 0 && gd < d.Di) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], bsc[j], bsh[j]);          // channels >= Ci: 0 * 1 + 0
                    }
                    const uint32_t s = gin % RING, use = gin / RING;
                    long long c0 = pr ? clock64() : 0;
                    mbar_wait_warp(a_empty + 8 * s, (use & 1) ^ 1, dbg, 0x100u + s, gin);
                    long long c1 = pr ? clock64() : 0;
                    pw0 += c1 - c0;
                    uint4* Ab = As + (size_t)s * SLOT_U4 + (size_t)chunk * CHS + slot;
                    if (NS == 3) {
                        uint4 o[3];
                        split8_trunc3(v, o);
#pragma unroll
                        for (int s2 = 0; s2 < 3; ++s2) Ab[(size_t)s2 * 2 * CHS] = o[s2];
                    } else {
                        Ab[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full + 8 * s);
                    if (pr) pwk += clock64() - c1;
                }
                continue;
            }
            // (fallback when the source cannot be described by a tensor map: channel stride not a multiple of 16 bytes)
            // register ring of PF planes: plane ip lives in entry ip % PF; after it is consumed its entry is reloaded with
            // plane ip + PF, so PF - 1 .. PF planes of loads are in flight while one is split and stored
            float4 pa[PF], pb[PF];
            bool pin[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                pin[u] = false;
                if (u < nplanes) pin[u] = load_plane(u, pa[u], pb[u]);
            }
#pragma unroll 1
            for (int ip0 = 0; ip0 < nplanes; ip0 += PF) {
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int ip = ip0 + u;
                    if (ip < nplanes) {
                        const uint32_t s = gin % RING, use = gin / RING;
                        long long c0 = pr ? clock64() : 0;
                        mbar_wait_warp(a_empty + 8 * s, (use & 1) ^ 1, dbg, 0x100u + s, gin);   // the output planes that read this slot RING planes ago are done
                        long long c1 = pr ? clock64() : 0;
                        pw0 += c1 - c0;
                        float v[8] = {pa[u].x, pa[u].y, pa[u].z, pa[u].w, pb[u].x, pb[u].y, pb[u].z, pb[u].w};
                        if (pin[u] && scale) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], bsc[j], bsh[j]);      // channels >= Ci: 0 * 1 + 0
                        }
                        if (ip + PF < nplanes) pin[u] = load_plane(ip + PF, pa[u], pb[u]);
                        uint4* Ab = As + (size_t)s * SLOT_U4 + (size_t)chunk * CHS + slot;
                        if (NS == 3) {
                            uint4 o[3];
                            split8_trunc3(v, o);
#pragma unroll
                            for (int s2 = 0; s2 < 3; ++s2) Ab[(size_t)s2 * 2 * CHS] = o[s2];
                        } else {
                            Ab[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                        }
                        fence_async_smem();                          // generic-proxy writes -> visible to the tensor core
                        __syncwarp();
                        if (lane == 0) mbar_arrive(a_full + 8 * s);
                        ++gin;
                        if (pr) pwk += clock64() - c1;
                    }
                }
            }
        }
        if (pr && st == 0) { prof[4] = pw0; prof[5] = pwk; }
    } else if (warp >= NEPI_W + NSTG_W) {
        // =================================================================== MMA issue: the whole issuer warp runs the loop with
        // warp-uniform values, one elected lane issues (umma_bf16_elect: inside an `if (lane == 0)` region every tcgen05.mma is
        // wrapped in an ELECT / R2UR / BRA.U.ANY loop of ~17 instructions)
        const int iss = warp_uniform(warp - (NEPI_W + NSTG_W));
        if (iss < NACC) {
            const uint32_t a_base = smem_u32(As), b_base = smem_u32(Bs);
            const uint64_t db0 = umma_desc(b_base, BROWS * 16, 128);
            uint32_t q = 0, gbase = 0;                           // global output-plane / input-plane counters
            int nq = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                const Item3 it = item_of(item);
                for (int od = 0; od < it.L; ++od, ++q) {
                    if ((int)(q % NACC) != iss) continue;
                    const uint32_t g2 = gbase + od + 2;          // the last of the three input planes this output plane reads
                    long long c0 = pr ? clock64() : 0;
                    mbar_wait3(a_full + 8 * (g2 % RING), (g2 / RING) & 1, dbg, 0x200u + (g2 % RING) + 16u * iss, g2);
                    long long c1 = pr ? clock64() : 0;
                    const uint32_t acc = q % NACC;
                    mbar_wait3(t_empty + 8 * acc, ((q / NACC) & 1) ^ 1, dbg, 0x300u + acc, q);      // the epilogue drained this accumulator
                    long long c2 = pr ? clock64() : 0;
                    pw0 += c1 - c0; pw1 += c2 - c1;
                    tc_fence_after();
                    const uint32_t dk = tmem_base + acc * ACOLS;
#pragma unroll 1
                    for (int kd = 0; kd < 3; ++kd) {
                        const uint32_t s = (gbase + od + kd) % RING;
                        const uint32_t a_slot = a_base + s * (SLOT_U4 * 16);
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const uint64_t db = db0 + (uint64_t)((kd * 3 + kh) * 2 * BROWS);
                            const uint64_t da = umma_desc(a_slot + (uint32_t)(kh * TW * 16), CHS * 16, 128);
                            const uint32_t nfirst = (kd | kh) != 0;
                            //   a1 x [w1|w2|w3] -> [main | cA | cB];  a2 x [w1|w2] -> [cA | cB];  a3 x [w1] -> cA
                            umma_bf16_elect(dk, da, db, umma_idesc_bf16(NTOT), nfirst);
                            if (NS == 3) {
                                umma_bf16_elect(dk + (uint32_t)TS, da + (uint64_t)(1 * 2 * CHS), db, umma_idesc_bf16(2 * TS), 1u);
                                umma_bf16_elect(dk + (uint32_t)TS, da + (uint64_t)(2 * 2 * CHS), db, umma_idesc_bf16(TS), 1u);
                            }
                        }
                    }
                    umma_commit_elect(t_full + 8 * (q % NTF));         // accumulator complete -> the epilogue group of this plane
                    // release the input planes: slot of plane ip is free once its (up to) three reading output planes are done;
                    // planes at the ends of a segment have fewer readers, their last reader arrives for the missing ones
                    for (int kd = 0; kd < 3; ++kd) {
                        const int ip = od + kd;
                        const int last = ip < it.L - 1 ? ip : it.L - 1, first = ip - 2 > 0 ? ip - 2 : 0;
                        const int narr = 1 + (od == last ? 3 - (last - first + 1) : 0);
                        const uint32_t s = (gbase + ip) % RING;
                        for (int a = 0; a < narr; ++a) umma_commit_elect(a_empty + 8 * s);
                    }
                    if (pr) pwk += clock64() - c2;
                    ++nq;
                }
                gbase += it.L + 2;
            }
            if (pr && iss == 0 && lane == 0) { prof[0] = pw0; prof[1] = pw1; prof[2] = pwk; prof[3] = nq; }
        }
    } else {
        // =================================================================== epilogue warps (TMEM lane quarter = warp % 4)
        const int grp = warp >> 2, q4 = warp & 3;
        const int r = q4 * 32 + lane;                            // GEMM row = (h, w'') of the tile
        const int hy = r / TW, wx = r % TW;
        const bool has_bias = bias != nullptr && fin;
        uint32_t q = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item3 it = item_of(item);
            const int oh = it.oh0 + hy, ow = it.ow0 + wx;
            const bool valid_hw = wx < TWV && oh < d.Ho && ow < d.Wo;
            for (int od = 0; od < it.L; ++od, ++q) {
                if ((int)(q % NEPI_G) != grp) continue;
                const uint32_t acc = q % NACC;
                long long c0 = pr ? clock64() : 0;
                mbar_wait_warp(t_full + 8 * (q % NTF), (q / NTF) & 1, dbg, 0x400u + acc + 16u * grp, q);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                tc_fence_after();
                const uint32_t ta = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * ACOLS;
                const int odg = it.od_lo + od;
                float* yp = dst + ((((int64_t)it.n * d.Do + odg) * d.Ho + oh) * d.Wo + ow) * d.ldo;
                const bool vecy = (d.ldo % 4 == 0) && (d.Co % 8 == 0);
                // Eight output channels at a time (24 + 16 live registers: with every role's registers capped at 80 by the CTA's 21 warps
                // the 16-wide form spilled, and the epilogue then waited for its own spill traffic).
#pragma unroll
                for (int c0 = 0; c0 < COP; c0 += 8) {
                    float out[8];
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        float s[8];
                        if (NS == 3) {
                            uint32_t m[8], a[8], b[8];
                            tmem_ld_issue<8>(ta + kw * COP + c0, m);                  // main
                            tmem_ld_issue<8>(ta + TS + kw * COP + c0, a);             // cA
                            tmem_ld_issue<8>(ta + 2 * TS + kw * COP + c0, b);         // cB
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                s[j] = (__uint_as_float(b[j]) + __uint_as_float(a[j])) + __uint_as_float(m[j]);   // corrections first
                        } else {
                            uint32_t m[8];
                            tmem_ld_issue<8>(ta + kw * COP + c0, m);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 8; ++j) s[j] = __uint_as_float(m[j]);
                        }
                        if (kw == 2 && c0 + 8 >= COP) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(t_empty + 8 * acc);  // this accumulator may be overwritten (one arrival per warp)
                        }
                        // out[w] = D0[w] + D1[w + 1] + D2[w + 2]: rows w'' + kw of the same h are lanes + kw of this warp
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float t = (kw == 0) ? s[j] : __shfl_down_sync(0xffffffffu, s[j], kw);
                            out[j] = (kw == 0) ? t : out[j] + t;
                        }
                    }
                    if (valid_hw && c0 < d.Co) {
                        if (accum) {                              // raw sums of the earlier input-channel passes
                            if (vecy) {
                                const float4 o0 = reinterpret_cast<const float4*>(yp + c0)[0], o1 = reinterpret_cast<const float4*>(yp + c0)[1];
                                out[0] += o0.x; out[1] += o0.y; out[2] += o0.z; out[3] += o0.w;
                                out[4] += o1.x; out[5] += o1.y; out[6] += o1.z; out[7] += o1.w;
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (c0 + j < d.Co) out[j] += yp[c0 + j];
                            }
                        }
                        if (fin) {
                            float bch[8];                         // (read per use: L1-resident, and no registers held across the plane loop)
#pragma unroll
                            for (int j = 0; j < 8; ++j) bch[j] = (has_bias && c0 + j < d.Co) ? __ldg(bias + c0 + j) : 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) out[j] += bch[j];
                            // one branch on the (uniform) activation per octet, never per element
                            if (d.act == SP_ACT_ELU) {
                                sp_elu_n<8>(out, d.alpha);
                            } else if (d.act == SP_ACT_LEAKY) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) out[j] = fmaxf(out[j], 0.f) + d.alpha * fminf(out[j], 0.f);
                            } else if (d.act == SP_ACT_SIGMOID) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) out[j] = 1.f / (1.f + expf(-out[j]));
                            }
                        }
                        if (vecy) {
                            reinterpret_cast<float4*>(yp + c0)[0] = make_float4(out[0], out[1], out[2], out[3]);
                            reinterpret_cast<float4*>(yp + c0)[1] = make_float4(out[4], out[5], out[6], out[7]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (c0 + j < d.Co) yp[c0 + j] = out[j];
                        }
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
}

}  // namespace sp_tc3

// ---- TMA: tensor map of the fp32 NDHWC source --------------------------------------------------------------------------
typedef CUresult (*SpEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline SpEncodeTiledFn sp_tma_encoder() {
    static SpEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {       // the driver entry point through the runtime: no link-time dependency on libcuda
        tried = true;
        const char* off = getenv("SP_TC3_NO_TMA");
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (!(off && off[0] == '1') && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<SpEncodeTiledFn>(p);
    }
    return fn;
}
// 5-D map (c, w, h, d, n) of `src` (ci channels used of ldi per voxel) with the box [16 c][16 w][10 h][1][1]; false when the
// source cannot be described (voxel stride not a multiple of 16 bytes, misaligned base): the kernel then stages with plain loads
static inline bool sp_tc3_make_tmap(const SpConvDesc* d, const float* src, CUtensorMap* tm) {
    SpEncodeTiledFn enc = sp_tma_encoder();
    if (!enc || d->ldi % 4 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0) return false;
    const cuuint64_t dims[5] = {(cuuint64_t)d->Ci, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->Di, (cuuint64_t)d->N};
    const cuuint64_t vs = (cuuint64_t)d->ldi * 4;
    const cuuint64_t strides[4] = {vs, vs * d->Wi, vs * d->Wi * d->Hi, vs * d->Wi * d->Hi * d->Di};
    const cuuint32_t box[5] = {16, (cuuint32_t)sp_tc3::TW, (cuuint32_t)sp_tc3::IH, 1, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(src), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline size_t sp_tc3_wimg_u4(int cop, int ns) { return (size_t)9 * 2 * ns * ((3 * cop + 15) / 16 * 16); }

// one image per (output slice, input-channel pass), image (s, p) at index s * passes + p
static inline int sp_tc3_pack_launch(const SpConvDesc* d, int transposed, int ns, int cop, const float* w, void* img, cudaStream_t st,
                                     int passes = 1, int nslices = 1) {
    const int ts = (3 * cop + 15) / 16 * 16;
    const int total = 9 * 2 * ts;
    const int blocks = (total + 255) / 256;
    const size_t img_u4 = sp_tc3_wimg_u4(cop, ns);
    for (int sl = 0; sl < nslices; ++sl)
        for (int p = 0; p < passes; ++p) {
            uint4* ip = (uint4*)img + (size_t)(sl * passes + p) * img_u4;
            if (ns == 3) sp_tc3::pack_wimg3_kernel<3><<<blocks, 256, 0, st>>>(w, d->Co, d->Ci, transposed, cop, ts, 16 * p, 16 * sl, ip);
            else sp_tc3::pack_wimg3_kernel<1><<<blocks, 256, 0, st>>>(w, d->Co, d->Ci, transposed, cop, ts, 16 * p, 16 * sl, ip);
            SP_LAUNCH_OK("pack_wimg3_kernel");
        }
    return 0;
}

template <int COP, int NS>
static inline int sp_tc3_corr_launch_t(const SpConvDesc* d, int nPerG, const float* src, const uint4* wimg, const float* bias,
                                       const float* scale, const float* shift, int sstride, int accum, int fin, float* dst,
                                       cudaStream_t st, long long* prof) {
    using namespace sp_tc3;
    const int tiles_w = (d->Wo + TWV - 1) / TWV, tiles_h = (d->Ho + TH - 1) / TH;
    const int64_t columns = (int64_t)tiles_w * tiles_h * d->N;
    // depth segments: whole columns when there are enough of them to balance the SMs, else split (2 extra halo planes per segment)
    const int sms = sp_num_sms();
    int nseg = 1;
    if (columns < 8LL * sms) {
        nseg = (int)((8LL * sms + columns - 1) / columns);
        const int max_seg = d->Do / 6 > 1 ? d->Do / 6 : 1;
        if (nseg > max_seg) nseg = max_seg;
    }
    const int seg_len = (d->Do + nseg - 1) / nseg;
    nseg = (d->Do + seg_len - 1) / seg_len;
    const int64_t total = columns * nseg;
    SP_REQUIRE(total < (1LL << 31), "tc3 corr: too many work items");
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(corr3_tc3_kernel<COP, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tc3<COP, NS>::SMEM));
        attr = true;
    }
    int grid = sms;
    static int dbg = -1, dbg_grid = 0;                   // diagnostics: SP_TC3_DEBUG=1 (timeout records), SP_TC3_GRID=n (CTAs)
    if (dbg < 0) {
        const char* e = getenv("SP_TC3_DEBUG");
        dbg = (e && e[0] == '1') ? 1 : 0;
        const char* g = getenv("SP_TC3_GRID");
        dbg_grid = g ? atoi(g) : 0;
    }
    if (dbg_grid > 0 && dbg_grid < grid) grid = dbg_grid;
    if (grid > total) grid = (int)total;
    alignas(64) CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    const int use_tma = sp_tc3_make_tmap(d, src, &tm) ? 1 : 0;
    corr3_tc3_kernel<COP, NS><<<grid, NTHREADS3, Tc3<COP, NS>::SMEM, st>>>(*d, nPerG, tiles_w, tiles_h, nseg, seg_len, (int)total, src, wimg,
                                                                           bias, scale, shift, sstride, accum, fin, dst, prof, dbg, tm, use_tma);
    SP_LAUNCH_OK("corr3_tc3_kernel");
    return 0;
}

// d->Ci <= 16, d->Co <= 24: one launch.  Wider layers: slices of 16 output channels (each its own launch into a channel slice of
// dst), input channels in passes of 16 whose raw sums accumulate in dst.  ns = 3: fp32-grade split arithmetic, 1: bf16 mode.
static inline int sp_tc3_corr_launch(const SpConvDesc* d, int nPerG, int ns, const float* src, const uint4* wimg, const float* bias,
                                     const float* scale, const float* shift, float* dst, cudaStream_t st, long long* prof = nullptr) {
    const int cop = (d->Co > 16 && d->Co <= 24) ? 24 : 16;
    const int nsl = (d->Co <= 24) ? 1 : (d->Co + 15) / 16;
    const int npass = (d->Ci + 15) / 16;
    const size_t img_u4 = sp_tc3_wimg_u4(cop, ns);
    for (int sl = 0; sl < nsl; ++sl)
        for (int p = 0; p < npass; ++p) {
            SpConvDesc s = *d;
            s.Ci = (d->Ci - 16 * p < 16) ? d->Ci - 16 * p : 16;
            if (nsl > 1) s.Co = (d->Co - 16 * sl < 16) ? d->Co - 16 * sl : 16;
            const float* sp = src + 16 * p;
            const float* scp = scale ? scale + 16 * p : nullptr;
            const float* shp = shift ? shift + 16 * p : nullptr;
            const float* bp = bias ? bias + 16 * sl : nullptr;
            float* dp = dst + 16 * sl;
            const uint4* ip = wimg + (size_t)(sl * npass + p) * img_u4;
            const int accum = p > 0, fin = p == npass - 1;
            int e;
            if (cop == 24) e = ns == 3 ? sp_tc3_corr_launch_t<24, 3>(&s, nPerG, sp, ip, bp, scp, shp, d->Ci, accum, fin, dp, st, prof)
                                       : sp_tc3_corr_launch_t<24, 1>(&s, nPerG, sp, ip, bp, scp, shp, d->Ci, accum, fin, dp, st, prof);
            else e = ns == 3 ? sp_tc3_corr_launch_t<16, 3>(&s, nPerG, sp, ip, bp, scp, shp, d->Ci, accum, fin, dp, st, prof)
                             : sp_tc3_corr_launch_t<16, 1>(&s, nPerG, sp, ip, bp, scp, shp, d->Ci, accum, fin, dp, st, prof);
            if (e) return e;
        }
    return 0;
}
