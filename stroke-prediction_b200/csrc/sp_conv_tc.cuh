// sp_conv_tc.cuh — tcgen05 / TMEM implicit-GEMM tier for the 3x3x3 stride-1 correlations (16..64 channels) that dominate
// the CAE / U-Net step (Cae3D.py:44,52,55,186-211; Unet3D.py:19,22) and, through flipped taps, their stride-1 dgrads.
//
// Why tensor cores for an "fp32" layer: a C->C 3x3x3 conv has 27*C/4 FLOP per byte (>= 108 for C >= 16), far above the
// FP32-FFMA ridge (~11 FLOP/B), so the FFMA tier (sp_conv_tiled.cuh) is compute-bound at <= 11 % of the HBM roof.  Here
// the MACs run on the 5th-gen tensor cores with SPLIT operands: every fp32 value v is staged as NS bf16 terms
// v = v1 + v2 (+ v3) (vi = bf16(v - v1 - .. - v(i-1))) and the products of order <= NS-1 are accumulated in fp32 in TMEM:
//      NS = 2:  a1*w1 + a1*w2 + a2*w1                        (relative product error ~ 2^-17)
//      NS = 3:  a1*w1 + a1*w2 + a2*w1 + a1*w3 + a2*w2 + a3*w1 (~ 2^-24: fp32-equivalent)
// The weight terms are stacked along N (B = [w1|w2|w3], one MMA per activation term: a1 x N=NS*Co, a2 x N=(NS-1)*Co, ..)
// so the A operand, whose shared-memory reads bound the N=16..32 MMAs, is fetched once per term.
//
// Tile: one CTA = 8(w) x 16(h) x TD(d) output voxels x all output channels; GEMM view per depth plane:
//      D[128 voxels][N] += A[128 voxels][K = 16 ci] * B[N][K]        for each of the 27 taps and each 16-channel K step.
// The BN-applied, zero-padded input halo tile (10 x 18 x (TD+2) voxels) is staged ONCE as "channel-chunk planes":
// plane (term s, chunk c) holds one 16-byte slot (8 bf16 channels) per halo voxel, slot = (dz*18 + hy)*10 + wx.  That is
// exactly the canonical K-major SWIZZLE_NONE UMMA layout (8 rows x 16 B core matrices): rows 0..7 of a core matrix are 8
// consecutive w voxels, the next core matrix along M is the next h row (SBO = 10 slots), the next along K the next
// channel plane (LBO = plane stride) — so the A operand of tap (kd,kh,kw) is the SAME buffer with the descriptor start
// address advanced by ((kd*18 + kh)*10 + kw) slots: no im2col, no per-tap re-staging.
#pragma once
#include <cuda_bf16.h>
#include "sp_common.cuh"

namespace sp_tc {

constexpr int TWO = 8, THO = 16;            // output tile (w, h); 128 voxels = UMMA M
constexpr int IWP = TWO + 2, IHP = THO + 2; // halo tile extents (slots per row, rows per plane)
constexpr int NTHREADS = 256;

template <int TD>
__host__ __device__ constexpr int slots() { return (TD + 2) * IHP * IWP; }

// shared memory image of the packed weights: [27 taps][CIP/8 chunks][NS*COP rows] x 16 B (8 bf16 input channels)
template <int CIP, int COP, int NS>
__host__ __device__ constexpr int wimg_u4() { return 27 * (CIP / 8) * NS * COP; }

template <int CIP, int COP, int NS, int TD>
__host__ __device__ constexpr size_t smem_bytes() {
    return (size_t)NS * (CIP / 8) * slots<TD>() * 16 + (size_t)wimg_u4<CIP, COP, NS>() * 16 + 64;
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell):
// bits [0,14) start >> 4, [16,30) leading-dimension byte offset >> 4 (K-direction core-matrix stride),
// [32,46) stride byte offset >> 4 (M/N-direction core-matrix stride), [46,48) version, [61,64) layout type (0).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major, M = 128.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// Warp-converged forms: the WHOLE warp executes the call with warp-uniform arguments, one elected lane issues.  Inside an
// `if (lane == 0)` region the compiler cannot prove uniformity and wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST /
// BRA.U.ANY loop (~17 SASS instructions and ~215 cycles per MMA and issuing thread, measured in the weight-gradient kernels);
// here the descriptors live in uniform registers.  warp_uniform() makes a per-warp value uniform for the compiler.
__device__ __forceinline__ int warp_uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
// bounded spin: a wrong descriptor must fail the launch (trap), never hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot_in_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread t of the warp reads TMEM lane 32*(warp%4)+t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__host__ __device__ constexpr int tmem_cols_pow2(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// ---- split helpers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);      // .x = lo (low 16 bits), .y = hi
    return *reinterpret_cast<uint32_t*>(&t);
}
// v[8] -> NS uint4 of bf16 terms (term s holds bf16 of the residual after terms < s)
template <int NS>
__device__ __forceinline__ void split8(const float* v, uint4* out) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat16 b0 = __float2bfloat16_rn(r[2 * i]), b1 = __float2bfloat16_rn(r[2 * i + 1]);
            r[2 * i] -= __bfloat162float(b0);
            r[2 * i + 1] -= __bfloat162float(b1);
            w[i] = (uint32_t)__bfloat16_as_ushort(b0) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
        }
        out[s] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- weight image -----------------------------------------------------------------------------------------------------
// img[((tap*KCH + chunk)*NTOT + s*COP + co)] = 8 bf16 {term s of Wsrc(co, chunk*8 + j, tap)}, j = 0..7.
// transposed == 0 (sp_corr):  Wsrc(co, ci, tap) = w[(co*Ci + ci)*27 + tap]            GEMM N = co, K = ci
// transposed == 1 (sp_corrT): GEMM N = conv-ci, K = conv-co, taps flipped: Wsrc(n, k, tap) = w[(k*Ci + n)*27 + 26 - tap]
template <int NS>
__global__ void pack_wimg_kernel(const float* __restrict__ w, int Co, int Ci, int transposed, int CIP, int COP, int k0, int n0, uint4* __restrict__ img) {
    const int KCH = CIP / 8, NTOT = NS * COP;
    const int total = 27 * KCH * COP;
    const int Nn = transposed ? Ci : Co, Kk = transposed ? Co : Ci;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int nl = i % COP;
        const int n = n0 + nl;                        // n0: first GEMM-N channel of this output slice
        const int chunk = (i / COP) % KCH;
        const int tap = i / (COP * KCH);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + chunk * 8 + j;          // k0: first GEMM-K channel of this input-channel pass
            float x = 0.f;
            if (n < Nn && k < Kk)
                x = transposed ? w[((int64_t)k * Ci + n) * 27 + (26 - tap)] : w[((int64_t)n * Ci + k) * 27 + tap];
            v[j] = x;
        }
        uint4 o[NS];
        split8<NS>(v, o);
#pragma unroll
        for (int s = 0; s < NS; ++s) img[(tap * KCH + chunk) * NTOT + s * COP + nl] = o[s];
    }
}

// ---- the kernel -------------------------------------------------------------------------------------------------------
// d: correlation geometry (k = 3, s = 1) with d.Ci <= CIP, d.Co <= COP.  src fp32 NDHWC (ldi), dst fp32 NDHWC (ldo).
template <int CIP, int COP, int NS, int TD>
__global__ void __launch_bounds__(NTHREADS, 2)
corr3_tc_kernel(SpConvDesc d, int nPerG, int tiles_w, int tiles_h, int tiles_d, int total_tiles,
                const float* __restrict__ src, const uint4* __restrict__ wimg, const float* __restrict__ bias,
                const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ dst,
                long long* __restrict__ prof) {
    constexpr int KCH = CIP / 8;                 // 16-byte channel chunks per voxel and term
    constexpr int KSTEPS = CIP / 16;             // UMMA K = 16 bf16
    constexpr int SLOTS = slots<TD>();
    constexpr int PLANE_B = SLOTS * 16;          // bytes per (term, chunk) plane
    constexpr int NTOT = NS * COP;
    constexpr int WIMG = wimg_u4<CIP, COP, NS>();
    constexpr int TCOLS = tmem_cols_pow2(TD * NTOT);
    static_assert(TD * NTOT <= 512, "accumulators exceed TMEM");
    static_assert(COP % 16 == 0 && CIP % 16 == 0, "UMMA M=128 needs N % 16 == 0; K step is 16");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* As = reinterpret_cast<uint4*>(smem_raw);                         // [NS][KCH][SLOTS]
    uint4* Bs = As + (size_t)NS * KCH * SLOTS;                              // weight image
    uint64_t* bar = reinterpret_cast<uint64_t*>(Bs + WIMG);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: weights -> smem, mbarrier, TMEM
    for (int i = tid; i < WIMG; i += NTHREADS) Bs[i] = wimg[i];
    if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<TCOLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_base = smem_u32(As), b_base = smem_u32(Bs), bar_a = smem_u32(bar);
    uint32_t phase = 0;

    const bool vec = (d.ldi % 4 == 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int t = tile;
        const int tw = t % tiles_w; t /= tiles_w;
        const int th_ = t % tiles_h; t /= tiles_h;
        const int td_ = t % tiles_d;
        const int n = t / tiles_d;
        const int ow0 = tw * TWO, oh0 = th_ * THO, od0 = td_ * TD;
        const int id0 = od0 - d.pd, ih0 = oh0 - d.ph, iw0 = ow0 - d.pw;
        const int g = n / nPerG;
        const float* srcn = src + (int64_t)n * d.Di * d.Hi * d.Wi * d.ldi;
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        const bool pr = (prof != nullptr) && (blockIdx.x == 0) && (tid == 0);
        if (pr) t0 = clock64();

        // ---- stage the halo tile: BN applied, zero padding written as zeros, split into NS bf16 terms.
        // item = (16 consecutive slots) x chunk: half-warps read 32-byte channel chunks of consecutive voxels and write
        // consecutive 16-byte slots of one plane (conflict-free STS.128).
        constexpr int SLOT_GROUPS = (SLOTS + 15) / 16;
        for (int it = tid; it < SLOT_GROUPS * KCH * 16; it += NTHREADS) {
            const int sub = it & 15;
            const int chunk = (it >> 4) % KCH;
            const int slot = ((it >> 4) / KCH) * 16 + sub;
            if (slot >= SLOTS) continue;
            const int wx = slot % IWP;
            const int hy = (slot / IWP) % IHP;
            const int dz = slot / (IWP * IHP);
            const int gd = id0 + dz, gh = ih0 + hy, gw = iw0 + wx;
            const int c = chunk * 8;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            if (gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi && c < d.Ci) {
                const float* p = srcn + (((int64_t)gd * d.Hi + gh) * d.Wi + gw) * d.ldi + c;
                if (vec && c + 8 <= d.Ci) {
                    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c + j < d.Ci) v[j] = p[j];
                }
                if (scale) {
                    const float* sc = scale + (int64_t)g * d.Ci + c;
                    const float* sh = shift + (int64_t)g * d.Ci + c;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c + j < d.Ci) v[j] = fmaf(v[j], sc[j], sh[j]);
                }
            }
            uint4 o[NS];
            split8<NS>(v, o);
#pragma unroll
            for (int s = 0; s < NS; ++s) As[((size_t)s * KCH + chunk) * SLOTS + slot] = o[s];
        }
        fence_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (pr) t1 = clock64();

        // ---- MMA issue: one thread
        if (tid == 0) {
            tc_fence_after();
#pragma unroll 1
            for (int p = 0; p < TD; ++p) {
                const uint32_t dcol = tmem_base + (uint32_t)(p * NTOT);
                uint32_t acc = 0;
#pragma unroll 1
                for (int tap = 0; tap < 27; ++tap) {
                    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                    const uint32_t a_off = (uint32_t)((((p + kd) * IHP + kh) * IWP + kw) * 16);
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        const uint64_t db = umma_desc(b_base + (uint32_t)(((tap * KCH + 2 * ks) * NTOT) * 16), NTOT * 16, 128);
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            const uint64_t da = umma_desc(a_base + (uint32_t)((s * KCH + 2 * ks) * PLANE_B) + a_off, PLANE_B, IWP * 16);
                            umma_bf16(dcol, da, db, umma_idesc_bf16((NS - s) * COP), acc | (uint32_t)(s > 0));
                        }
                        acc = 1;
                    }
                }
            }
            umma_commit(bar_a);
        }
        if (pr) t2 = clock64();
        __syncwarp();
        mbar_wait(bar_a, phase);
        phase ^= 1;
        tc_fence_after();
        if (pr) t3 = clock64();

        // ---- epilogue: warps w and w+4 share TMEM lane quarter w; they split the depth planes
        {
            const int q = warp & 3;
            const int r = q * 32 + lane;                 // GEMM row = output voxel within the plane
            const int oh = oh0 + (r >> 3), ow = ow0 + (r & 7);
            const bool inb = (oh < d.Ho) && (ow < d.Wo);
            for (int p = (warp >> 2); p < TD; p += 2) {
                const int od = od0 + p;
                if (od >= d.Do) break;                   // warp-uniform
                float* yp = dst + ((((int64_t)n * d.Do + od) * d.Ho + oh) * d.Wo + ow) * d.ldo;
#pragma unroll
                for (int cb = 0; cb < COP; cb += 16) {
                    if (cb >= d.Co) break;
                    float accv[16], tv[16];
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * NTOT + cb);
                    // smallest terms first: (a1*w3) + (a1*w2 + a2*w2) + (a1*w1 + a2*w1 + a3*w1)
                    tmem_ld16(ta + (NS - 1) * COP, accv);
#pragma unroll
                    for (int s = NS - 2; s >= 0; --s) {
                        tmem_ld16(ta + s * COP, tv);
#pragma unroll
                        for (int j = 0; j < 16; ++j) accv[j] += tv[j];
                    }
                    if (inb) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int co = cb + j;
                            const float b = (bias && co < d.Co) ? bias[co] : 0.f;
                            accv[j] = sp_act_fwd(accv[j] + b, d.act, d.alpha);
                        }
                        if ((d.ldo % 4 == 0) && cb + 16 <= d.Co) {
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4)
                                reinterpret_cast<float4*>(yp + cb)[j4] = make_float4(accv[4 * j4], accv[4 * j4 + 1], accv[4 * j4 + 2], accv[4 * j4 + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (cb + j < d.Co) yp[cb + j] = accv[j];
                        }
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();             // TMEM drained and smem free before the next tile is staged / issued
        if (pr) {
            const long long t4 = clock64();
            prof[0] += t1 - t0; prof[1] += t2 - t1; prof[2] += t3 - t2; prof[3] += t4 - t3; prof[4] += 1;
        }
    }

    tc_fence_after();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem_base);
}

}  // namespace sp_tc

// ---- host side --------------------------------------------------------------------------------------------------------
// Mode of the tensor-core tier for the 8..96-channel 3x3x3 stride-1 correlations:
//   5 (default)  generation 3 (sp_conv_tc3.cuh): three-term split arithmetic, kw-stacked N, rolling depth window: per-layer forward
//                rel-L2 1.3e-7 against 2.6e-7 for an IEEE fp32 FFMA chain on the same data — fp32-grade, 4.5x the FFMA tier's speed
//   1            bf16 mode: generation-3 kernel with ONE bf16 term per operand (RN) and fp32 accumulation (per-layer rel-L2 ~2.5e-3)
//   4            generation 2: pipelined three-term kernel with split accumulators (sp_conv_tc2.cuh), same accuracy as 5
//   0            tier off: the exact-fp32 FFMA tier serves every layer
//   2 / 3        first-generation kernels (one accumulator per output, 2 / 3 bf16 terms: rel-L2 4.8e-6 / 1.4e-6 because the tensor
//                core's fp32 accumulator truncates); kept for A/B measurements
// Set at run time through sp_set_tc_terms(); the initial value may be given in the environment (SP_TC_TERMS=0|2|3|4).
static inline int& sp_tc_terms_ref() {
    static int v = -1;
    if (v < 0) {
        v = 5;
        const char* e = getenv("SP_TC_TERMS");
        if (e && e[0] >= '0' && e[0] <= '5') v = e[0] - '0';
    }
    return v;
}
static inline int sp_tc_terms() { return sp_tc_terms_ref(); }
static inline int sp_tc_image_terms() { return (sp_tc_terms() == 4 || sp_tc_terms() == 5) ? 3 : sp_tc_terms(); }   // bf16 terms of the weight image
// generation-3 kernel (sp_conv_tc3.cuh): mode 5 = three-term fp32-grade arithmetic, mode 1 = bf16 mode (one term, RN)
static inline bool sp_tc_gen3() { return sp_tc_terms() == 5 || sp_tc_terms() == 1; }

struct SpTcCfg { int cip, cop, td, passes, nslices; };   // pipelined kernel: input-channel passes of 16, output slices of 16 (Co > 24)

// layers served: 3x3x3 stride-1, <= 16 channels on both sides (the 28-deep 16->16 layers hold ~75 % of the conv FLOPs)
static inline bool sp_tc_corr_supported(const SpConvDesc* d, SpTcCfg* cfg) {
    if (d->k != 3 || d->s != 1 || sp_tc_terms() == 0) return false;
    // the pipelined kernel also serves wider layers: 17..24 output channels as one 24-wide pass, more as slices of 16 (each its
    // own launch into a channel slice of dst), and up to six input-channel passes of 16 whose raw sums accumulate in dst
    // (the 24-channel level of the CAE, every Block3x3x3 of the U-Net: Unet3D.py:19,22)
    const bool wide = sp_tc_terms() == 4 || sp_tc_gen3();
    const int comax = wide ? 96 : 16, cimax = wide ? 96 : 16;
    if (d->Ci > cimax || d->Co > comax || d->Ci < 8 || d->Co < 8) return false;   // narrower layers: FFMA tier (2- / 8-wide passes)
    const int64_t ov = (int64_t)d->Do * d->Ho * d->Wo;
    if (ov < 4096 || d->Wo < 8 || d->Ho < 16) return false;      // (the CAE's 32-channel level: 7x27x27 per sample)
    if (d->Co > 24 && d->ldo % 4 != 0) return false;
    if (cfg) {
        cfg->cip = 16; cfg->td = 4;
        cfg->cop = (d->Co > 16 && d->Co <= 24) ? 24 : 16;
        cfg->nslices = (d->Co <= 24) ? 1 : (d->Co + 15) / 16;
        cfg->passes = (d->Ci + 15) / 16;
    }
    return true;
}

static inline size_t sp_tc_wimg_bytes(int cip, int cop, int ns) { return (size_t)27 * (cip / 8) * ns * cop * 16; }

template <int CIP, int COP, int NS, int TD>
static inline int sp_tc_corr_launch_t(const SpConvDesc* d, int nPerG, const float* src, const uint4* wimg, const float* bias,
                                      const float* scale, const float* shift, float* dst, cudaStream_t st,
                                      long long* prof = nullptr) {
    using namespace sp_tc;
    const int tiles_w = (d->Wo + TWO - 1) / TWO, tiles_h = (d->Ho + THO - 1) / THO, tiles_d = (d->Do + TD - 1) / TD;
    const int64_t total = (int64_t)tiles_w * tiles_h * tiles_d * d->N;
    SP_REQUIRE(total < (1LL << 31), "tc corr: too many tiles");
    constexpr size_t smem = smem_bytes<CIP, COP, NS, TD>();
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(corr3_tc_kernel<CIP, COP, NS, TD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    int grid = 2 * sp_num_sms();
    if (grid > total) grid = (int)total;
    corr3_tc_kernel<CIP, COP, NS, TD><<<grid, NTHREADS, smem, st>>>(*d, nPerG, tiles_w, tiles_h, tiles_d, (int)total, src, wimg,
                                                                    bias, scale, shift, dst, prof);
    SP_LAUNCH_OK("corr3_tc_kernel");
    return 0;
}

// one image per (output slice, input-channel pass): GEMM-N channels [16 s, ..), GEMM-K channels [16 p, 16 p + 16); image
// (s, p) lies at index s * passes + p
static inline int sp_tc_pack_launch(const SpConvDesc* d, int transposed, int ns, int cip, int cop, const float* w, void* img,
                                    cudaStream_t st, int passes = 1, int nslices = 1) {
    const int total = 27 * (cip / 8) * cop;
    const int blocks = (total + 255) / 256;
    const size_t img_u4 = (size_t)27 * (cip / 8) * ns * cop;
    for (int sl = 0; sl < nslices; ++sl)
        for (int p = 0; p < passes; ++p) {
            uint4* ip = (uint4*)img + (size_t)(sl * passes + p) * img_u4;
            if (ns == 2) sp_tc::pack_wimg_kernel<2><<<blocks, 256, 0, st>>>(w, d->Co, d->Ci, transposed, cip, cop, 16 * p, 16 * sl, ip);
            else sp_tc::pack_wimg_kernel<3><<<blocks, 256, 0, st>>>(w, d->Co, d->Ci, transposed, cip, cop, 16 * p, 16 * sl, ip);
            SP_LAUNCH_OK("pack_wimg_kernel");
        }
    return 0;
}
