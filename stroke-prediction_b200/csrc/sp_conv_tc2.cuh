// sp_conv_tc2.cuh — warp-specialised, software-pipelined tcgen05 / TMEM correlation for the 16-channel 3x3x3 stride-1
// layers (Cae3D.py:44,208,211; Unet3D.py:22) and, through flipped taps, their dgrads.  Same operand layout and split-bf16
// arithmetic as sp_conv_tc.cuh (three bf16 terms per fp32 operand, products of order <= 2), with two changes:
//
//  * accuracy: the tensor core's fp32 accumulator truncates on every accumulation, so one accumulator fed by 81 MMAs is
//    ~10x noisier than an IEEE FFMA chain.  Here the leading products a1*w1 are accumulated in THREE accumulators (one per
//    kd: 9 accumulations each) and all correction products (a1*w2, a1*w3, a2*w1, a2*w2, a3*w1: 2^-8 .. 2^-16 of the result)
//    in separate columns; the epilogue adds the eleven partial sums in fp32 round-to-nearest, smallest first.
//    TMEM columns per output plane: ([main | c2 | c3] (the N = 48 MMA of term a1 writes 48 contiguous columns) + [c1 | c1']
//    (terms a2, a3)) x 3 kd = 240.  MMAs that accumulate into the same columns execute back to back (measured ~110 cycles
//    each when every MMA of a tile hits one accumulator set), so the issue order walks (kh, kw) outermost and kd innermost:
//    six independent accumulator chains per plane keep the tensor pipe busy.
//  * overlap: one persistent CTA per SM with three roles — warps 4-7 stage tile i+1 (global -> BN -> bf16 split -> shared
//    memory, double buffered), one thread of warp 8 issues the MMAs of tile i, warps 0-3 drain the TMEM accumulators of the
//    plane that just finished (bias + activation + store) while the MMAs of the other plane run.  mbarriers: a_full /
//    a_empty per shared-memory buffer, t_full / t_empty per TMEM plane.
#pragma once
#include "sp_conv_tc.cuh"

namespace sp_tc2 {

using namespace sp_tc;

constexpr int TD2 = 2;                               // output planes per tile
constexpr int NS2 = 3;
constexpr int CIP2 = 16;
constexpr int KCH2 = CIP2 / 8;
constexpr int SLOTS2 = slots<TD2>();                 // 720
// The two 16-byte K chunks of one operand row are fetched together: their planes must not be a multiple of 128 bytes apart
// (same banks -> the operand fetch serialises, measured 64 cycles per MMA).  Pad every chunk plane to an odd multiple of 64 B.
constexpr int PSLOTS2 = SLOTS2 + 4;                  // 724 slots -> 11584 B = 90.5 x 128
constexpr int PLANE_B2 = PSLOTS2 * 16;
constexpr int ABUF_U4 = NS2 * KCH2 * PSLOTS2;        // uint4 per A buffer
constexpr int NSTAGE = 192;                          // staging threads (warps 4..9)
constexpr int NTHREADS2 = 16 * 32;                   // 4 epilogue warps + 6 staging warps + 6 MMA warps

// COP = padded output channels of the GEMM: 16 (the 16-channel layers) or 24 (the 24-channel level of the CAE, whose
// 17..24 input channels run as TWO passes of 16 + 8 input channels, the second accumulating onto the first's raw sums)
template <int COP>
struct Tc2 {
    static constexpr int NTOT = NS2 * COP;                       // rows of the weight image per (tap, chunk)
    static constexpr int BROWS = NTOT + 4;                       // padded chunk stride in shared memory (832 / 1216 B)
    static constexpr int WIMG = wimg_u4<CIP2, COP, NS2>();       // global (unpadded) image
    static constexpr int WIMGS = 27 * KCH2 * BROWS;              // shared-memory image
    static constexpr int KDCOLS = 3 * COP;                       // TMEM columns per (plane, kd): [main | cA | cB]
    static constexpr int PCOLS = 3 * KDCOLS;                     // TMEM columns per plane (144 / 216)
    static constexpr size_t SMEM = (size_t)2 * ABUF_U4 * 16 + (size_t)WIMGS * 16 + 128;
};
static_assert(Tc2<24>::SMEM <= 227 * 1024 && 2 * Tc2<24>::PCOLS <= 512, "tc2: shared / tensor memory");

// Exact three-term bf16 split by TRUNCATION: t1 = v & 0xffff0000, r = v - t1, ... (24 significand bits = 3 x 8, every
// subtraction is exact).  Integer / FADD work only — the cvt-based split (sp_tc::split8) is bound by the 16-lane conversion
// pipe (48 conversions per 8 values).
__device__ __forceinline__ void split8_trunc3(const float* v, uint4* out) {
    uint32_t t1[8], t2[8], t3[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t u = __float_as_uint(v[i]);
        t1[i] = u & 0xffff0000u;
        const float r1 = v[i] - __uint_as_float(t1[i]);
        t2[i] = __float_as_uint(r1) & 0xffff0000u;
        const float r2 = r1 - __uint_as_float(t2[i]);
        t3[i] = __float_as_uint(r2);          // <= 8 significant bits left: already a bf16 value
    }
    // pack the high halves of (even, odd) channel pairs: low 16 bits = even channel
    auto pack = [](uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x7632); };
    out[0] = make_uint4(pack(t1[0], t1[1]), pack(t1[2], t1[3]), pack(t1[4], t1[5]), pack(t1[6], t1[7]));
    out[1] = make_uint4(pack(t2[0], t2[1]), pack(t2[2], t2[3]), pack(t2[4], t2[5]), pack(t2[6], t2[7]));
    out[2] = make_uint4(pack(t3[0], t3[1]), pack(t3[2], t3[3]), pack(t3[4], t3[5]), pack(t3[6], t3[7]));
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
// 80 consecutive columns (one kd block) = x64 + x16, one wait
__device__ __forceinline__ void tmem_ld80(uint32_t taddr, float* v) {
    uint32_t r[80];
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
    tmem_ld16_nowait(taddr + 64, r + 64);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 80; ++i) v[i] = __uint_as_float(r[i]);
}

// 48 consecutive columns (one kd block) = x32 + x16, one wait
__device__ __forceinline__ void tmem_ld48(uint32_t taddr, float* v) {
    uint32_t r[48];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    tmem_ld16_nowait(taddr + 32, r + 32);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 48; ++i) v[i] = __uint_as_float(r[i]);
}

// three 16-column blocks in flight, one wait
__device__ __forceinline__ void tmem_ld16x3(uint32_t a0, uint32_t a1, uint32_t a2, float* v0, float* v1, float* v2) {
    uint32_t r0[16], r1[16], r2[16];
    tmem_ld16_nowait(a0, r0);
    tmem_ld16_nowait(a1, r1);
    tmem_ld16_nowait(a2, r2);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        v0[i] = __uint_as_float(r0[i]);
        v1[i] = __uint_as_float(r1[i]);
        v2[i] = __uint_as_float(r2[i]);
    }
}

// NCOL consecutive columns in 8-column pieces, one wait
template <int NCOL>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, float* v) {
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

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// sstride: floats between the scale / shift rows of two statistics groups (the layer's full channel count when `src` is a
// channel slice); accum != 0: add the raw sums already in dst (second input-channel pass); fin == 0: store raw sums (no bias,
// no activation: first pass of two).
template <int COP>
__global__ void __launch_bounds__(NTHREADS2, 1)
corr3_tc_pipe_kernel(SpConvDesc d, int nPerG, int tiles_w, int tiles_h, int tiles_d, int total_tiles,
                     const float* __restrict__ src, const uint4* __restrict__ wimg, const float* __restrict__ bias,
                     const float* __restrict__ scale, const float* __restrict__ shift, int sstride, int accum, int fin,
                     float* __restrict__ dst, long long* __restrict__ prof, int dbg_terms) {
    using T = Tc2<COP>;
    constexpr int NTOT2 = T::NTOT, BROWS2 = T::BROWS, WIMG2 = T::WIMG, WIMG2S = T::WIMGS, KDCOLS = T::KDCOLS, PCOLS = T::PCOLS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const bool pr = (prof != nullptr) && (blockIdx.x == 0);
    long long pw0 = 0, pw1 = 0, pwk = 0;
    uint4* As = reinterpret_cast<uint4*>(smem_raw);                          // [2][NS][KCH][SLOTS]
    uint4* Bs = As + (size_t)2 * ABUF_U4;                                    // weight image
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + WIMG2S);                // a_full[2] a_empty[2] t_full[2] t_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < WIMG2; i += NTHREADS2) Bs[(i / NTOT2) * BROWS2 + (i % NTOT2)] = wimg[i];
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), NSTAGE); mbar_init(smem_u32(&bars[1]), NSTAGE);   // a_full: the staging threads
        mbar_init(smem_u32(&bars[2]), 3 * TD2); mbar_init(smem_u32(&bars[3]), 3 * TD2);   // a_empty: one tcgen05.commit per issuing warp
        mbar_init(smem_u32(&bars[4]), 3); mbar_init(smem_u32(&bars[5]), 3);         // t_full: the three kd issuers of the plane
        mbar_init(smem_u32(&bars[6]), 128); mbar_init(smem_u32(&bars[7]), 128);     // t_empty: the 128 epilogue threads
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[2]), t_full = smem_u32(&bars[4]), t_empty = smem_u32(&bars[6]);

    auto tile_origin = [&](int tile, int& n, int& od0, int& oh0, int& ow0) {
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th_ = t % tiles_h; t /= tiles_h;
        const int td_ = t % tiles_d;
        n = t / tiles_d;
        ow0 = tw * TWO; oh0 = th_ * THO; od0 = td_ * TD2;
    };

    if (warp >= 4 && warp < 10) {
        // =================================================================== staging warps
        const int st = tid - 128;
        const bool vec = (d.ldi % 4 == 0);
        static_assert(KCH2 == 2, "staging packs (slot, chunk) assuming two chunks");
        static_assert(NSTAGE % 32 == 0, "a staging thread must keep one channel chunk");
        const int my_chunk = (st >> 4) % KCH2;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1, use = it >> 1;
            long long c0 = pr ? clock64() : 0;
            mbar_wait(a_empty + 8 * buf, (use & 1) ^ 1);          // MMAs that read this buffer two tiles ago are done
            long long c1 = pr ? clock64() : 0;
            pw0 += c1 - c0;
            int n, od0, oh0, ow0;
            tile_origin(tile, n, od0, oh0, ow0);
            const int id0 = od0 - d.pd, ih0 = oh0 - d.ph, iw0 = ow0 - d.pw;
            const int g = n / nPerG;
            const float* srcn = src + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;
            uint4* Ab = As + (size_t)buf * ABUF_U4;
            // BatchNorm coefficients of this thread's channel chunk (fixed per thread: NSTAGE is a multiple of 32)
            float bsc[8], bsh[8];
            {
                const int c = my_chunk * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool okc = scale && (c + j < d.Ci);
                    bsc[j] = okc ? scale[(int64_t)g * sstride + c + j] : 1.f;
                    bsh[j] = okc ? shift[(int64_t)g * sstride + c + j] : 0.f;
                }
            }
            constexpr int SLOT_GROUPS = (SLOTS2 + 15) / 16;
            constexpr int NITEMS = SLOT_GROUPS * KCH2 * 16;
            constexpr int PER_R = 4;                                        // loads in flight per thread and round
            constexpr int ROUNDS = (NITEMS + NSTAGE * PER_R - 1) / (NSTAGE * PER_R);
#pragma unroll 1
            for (int rnd = 0; rnd < ROUNDS; ++rnd) {
                // phase 1: the global loads of this round in flight at once; phase 2: BN, split, store
                float4 ra[PER_R], rb[PER_R];
                int rslot[PER_R];
#pragma unroll
                for (int u = 0; u < PER_R; ++u) {
                    const int item = st + (rnd * PER_R + u) * NSTAGE;
                    const int sub = item & 15;
                    const int chunk = (item >> 4) % KCH2;
                    const int slot = ((item >> 4) / KCH2) * 16 + sub;
                    ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    rb[u] = ra[u];
                    rslot[u] = -1;
                    if (item < NITEMS && slot < SLOTS2) {
                        const int wx = slot % IWP;
                        const int hy = (slot / IWP) % IHP;
                        const int dz = slot / (IWP * IHP);
                        const int gd = id0 + dz, gh = ih0 + hy, gw = iw0 + wx;
                        const int c = chunk * 8;
                        rslot[u] = slot * 2 + chunk;              // valid item; bit 30 set below when the voxel is inside
                        if (gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi && c < d.Ci) {
                            rslot[u] |= 1 << 30;
                            const float* p = srcn + (((int64_t)gd * d.Hi + gh) * d.Wi + gw) * d.ldi + c;
                            if (vec && c + 8 <= d.Ci) {
                                ra[u] = *reinterpret_cast<const float4*>(p);
                                rb[u] = *reinterpret_cast<const float4*>(p + 4);
                            } else {
                                float e[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) e[j] = (c + j < d.Ci) ? p[j] : 0.f;
                                ra[u] = make_float4(e[0], e[1], e[2], e[3]);
                                rb[u] = make_float4(e[4], e[5], e[6], e[7]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < PER_R; ++u) {
                    if (rslot[u] < 0) continue;
                    const bool inside = (rslot[u] >> 30) & 1;
                    const int sc2 = rslot[u] & 0x3fffffff;
                    const int slot = sc2 >> 1, chunk = sc2 & 1;
                    const int c = chunk * 8;
                    float v[8] = {ra[u].x, ra[u].y, ra[u].z, ra[u].w, rb[u].x, rb[u].y, rb[u].z, rb[u].w};
                    if (inside && scale) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], bsc[j], bsh[j]);      // channels >= Ci: 0 * 1 + 0
                    }
                    (void)c;
                    uint4 o[NS2];
                    split8_trunc3(v, o);
#pragma unroll
                    for (int s2 = 0; s2 < NS2; ++s2) Ab[((size_t)s2 * KCH2 + chunk) * PSLOTS2 + slot] = o[s2];
                }
            }
            fence_async_smem();                                   // generic-proxy writes -> visible to the tensor core
            mbar_arrive(a_full + 8 * buf);
            if (pr) pwk += clock64() - c1;
        }
        if (pr && st == 0) { prof[4] = pw0; prof[5] = pwk; }
    } else if (warp >= 10) {
        // =================================================================== MMA issue: warp 10 + 3p + kd, one thread each
        // The MMAs of one issuing thread retire strictly one after the other (~80 cycles each whatever N is), MMAs of
        // different issuing WARPS overlap (lanes of one warp do not).  Warp 10 + 3p + kd owns the accumulator block (plane p,
        // kd): 27 MMAs per tile (9 (kh,kw) x 3 terms), all into columns no other issuer touches.
        // (warp-converged issue: the whole warp runs the loop with warp-uniform values, one elected lane issues — umma_bf16_elect)
        {
            const int wi = warp_uniform(warp - 10);
            const int p = wi / 3, kd = wi % 3;
            const uint32_t a_base0 = smem_u32(As), b_base = smem_u32(Bs);
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1, use = it >> 1;
                long long c0 = pr ? clock64() : 0;
                mbar_wait(a_full + 8 * buf, use & 1);
                if (pr) pw0 += clock64() - c0;
                tc_fence_after();
                const uint32_t a_base = a_base0 + (uint32_t)buf * (ABUF_U4 * 16);
                long long c2 = pr ? clock64() : 0;
                mbar_wait(t_empty + 8 * p, (it & 1) ^ 1);         // the epilogue drained this plane's accumulators
                long long c3 = pr ? clock64() : 0;
                pw1 += c3 - c2;
                tc_fence_after();
                const uint32_t dk = tmem_base + (uint32_t)(p * PCOLS + kd * KDCOLS);
                // descriptors differ only in their 14-bit start-address field (16-byte units): add offsets to a base
                const uint64_t da0 = umma_desc(a_base + (uint32_t)((p + kd) * IHP * IWP * 16), PLANE_B2, IWP * 16);
                const uint64_t db0 = umma_desc(b_base + (uint32_t)(kd * 9 * KCH2 * BROWS2 * 16), BROWS2 * 16, 128);
#pragma unroll 1
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const uint32_t first = (kh | kw) != 0;
                        const uint64_t db = db0 + (uint64_t)((kh * 3 + kw) * KCH2 * BROWS2);
                        const uint64_t da = da0 + (uint64_t)(kh * IWP + kw);
                        //   a1 x [w1|w2|w3] -> [main | cA | cB];  a2 x [w1|w2] -> [cA | cB];  a3 x [w1] -> cA
                        // (main holds only the nine leading products a1*w1 of this kd; every correction term, 2^-8 .. 2^-16 of
                        // the result, goes to the two correction blocks)
                        if (dbg_terms & 1) umma_bf16_elect(dk, da, db, umma_idesc_bf16(3 * COP), first);
                        if (dbg_terms & 2) umma_bf16_elect(dk + (uint32_t)COP, da + (uint64_t)(1 * KCH2 * PSLOTS2), db, umma_idesc_bf16(2 * COP), 1u);
                        if (dbg_terms & 4) umma_bf16_elect(dk + (uint32_t)COP, da + (uint64_t)(2 * KCH2 * PSLOTS2), db, umma_idesc_bf16(COP), 1u);
                    }
                }
                umma_commit_elect(t_full + 8 * p);                      // this lane's share of plane p is complete
                umma_commit_elect(a_empty + 8 * buf);                   // ... and it no longer reads this A buffer
                if (pr) pwk += clock64() - c3;
            }
            if (pr && warp == 10 && lane == 0) { prof[0] = pw0; prof[1] = pw1; prof[2] = pwk; prof[3] = it; }
        }
    } else {
        // =================================================================== epilogue warps 0..3 (TMEM lane quarter = warp)
        const int q = warp;
        const int r = q * 32 + lane;                              // GEMM row = output voxel within the plane
        float b16[COP];
#pragma unroll
        for (int j = 0; j < COP; ++j) b16[j] = (bias && fin && j < d.Co) ? bias[j] : 0.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            int n, od0, oh0, ow0;
            tile_origin(tile, n, od0, oh0, ow0);
            const int oh = oh0 + (r >> 3), ow = ow0 + (r & 7);
#pragma unroll 1
            for (int p = 0; p < TD2; ++p) {
                long long c0 = pr ? clock64() : 0;
                mbar_wait(t_full + 8 * p, it & 1);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                tc_fence_after();
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * PCOLS);
                // per kd: (cB + cA) + main — corrections first, then the leading sum; then the three kd
                float acc[COP];
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    float v[KDCOLS];
                    tmem_ld_n<KDCOLS>(ta + kd * KDCOLS, v);       // [main | cA | cB]
                    if (kd == 2) {
                        tc_fence_before();
                        mbar_arrive(t_empty + 8 * p);             // TMEM of this plane may be overwritten
                    }
#pragma unroll
                    for (int j = 0; j < COP; ++j) {
                        const float s_kd = (v[2 * COP + j] + v[COP + j]) + v[j];
                        acc[j] = (kd == 0) ? s_kd : acc[j] + s_kd;
                    }
                }
                const int od = od0 + p;
                if (od < d.Do && oh < d.Ho && ow < d.Wo) {
                    float* yp = dst + ((((int64_t)n * d.Do + od) * d.Ho + oh) * d.Wo + ow) * d.ldo;
                    const bool vecy = (d.ldo % 4 == 0) && d.Co == COP;
                    if (accum) {                                  // raw sums of the first input-channel pass
                        if (vecy) {
#pragma unroll
                            for (int j4 = 0; j4 < COP / 4; ++j4) {
                                const float4 o = reinterpret_cast<const float4*>(yp)[j4];
                                acc[4 * j4] += o.x; acc[4 * j4 + 1] += o.y; acc[4 * j4 + 2] += o.z; acc[4 * j4 + 3] += o.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < COP; ++j)
                                if (j < d.Co) acc[j] += yp[j];
                        }
                    }
                    if (fin) {
#pragma unroll
                        for (int j = 0; j < COP; ++j) acc[j] = sp_act_fwd(acc[j] + b16[j], d.act, d.alpha);
                    }
                    if (vecy) {
#pragma unroll
                        for (int j4 = 0; j4 < COP / 4; ++j4)
                            reinterpret_cast<float4*>(yp)[j4] = make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < COP; ++j)
                            if (j < d.Co) yp[j] = acc[j];
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

}  // namespace sp_tc2

template <int COP>
static inline int sp_tc2_corr_launch_t(const SpConvDesc* d, int nPerG, const float* src, const uint4* wimg, const float* bias,
                                       const float* scale, const float* shift, int sstride, int accum, int fin, float* dst,
                                       cudaStream_t st, long long* prof, int dbg_terms) {
    using namespace sp_tc2;
    const int tiles_w = (d->Wo + TWO - 1) / TWO, tiles_h = (d->Ho + THO - 1) / THO, tiles_d = (d->Do + TD2 - 1) / TD2;
    const int64_t total = (int64_t)tiles_w * tiles_h * tiles_d * d->N;
    SP_REQUIRE(total < (1LL << 31), "tc corr: too many tiles");
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(corr3_tc_pipe_kernel<COP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tc2<COP>::SMEM));
        attr = true;
    }
    int grid = sp_num_sms();
    if (grid > total) grid = (int)total;
    corr3_tc_pipe_kernel<COP><<<grid, NTHREADS2, Tc2<COP>::SMEM, st>>>(*d, nPerG, tiles_w, tiles_h, tiles_d, (int)total, src, wimg, bias, scale,
                                                                        shift, sstride, accum, fin, dst, prof, dbg_terms);
    SP_LAUNCH_OK("corr3_tc_pipe_kernel");
    return 0;
}

// d->Ci <= 16, d->Co <= 16: one launch.  Wider layers: output width 24 (17..24 channels) or slices of 16 output channels (more),
// input channels in passes of 16; the weight image of (slice s, pass p) is image s * npass + p of `wimg`.
static inline int sp_tc2_corr_launch(const SpConvDesc* d, int nPerG, const float* src, const uint4* wimg, const float* bias,
                                     const float* scale, const float* shift, float* dst, cudaStream_t st, long long* prof = nullptr, int dbg_terms = 7) {
    using namespace sp_tc2;
    if (d->Ci <= 16 && d->Co <= 16)
        return sp_tc2_corr_launch_t<16>(d, nPerG, src, wimg, bias, scale, shift, d->Ci, 0, 1, dst, st, prof, dbg_terms);
    const int cop = (d->Co > 16 && d->Co <= 24) ? 24 : 16;
    const int nsl = (d->Co <= 24) ? 1 : (d->Co + 15) / 16;
    const int npass = (d->Ci + 15) / 16;
    const size_t img_u4 = (size_t)27 * KCH2 * NS2 * cop;
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
            const int e = (cop == 24)
                ? sp_tc2_corr_launch_t<24>(&s, nPerG, sp, ip, bp, scp, shp, d->Ci, p > 0, p == npass - 1, dp, st, prof, dbg_terms)
                : sp_tc2_corr_launch_t<16>(&s, nPerG, sp, ip, bp, scp, shp, d->Ci, p > 0, p == npass - 1, dp, st, prof, dbg_terms);
            if (e) return e;
        }
    return 0;
}
