// sp_wgrad_tc.cuh — tcgen05 / TMEM weight gradient of the 16-channel 3x3x3 stride-1 layers (Cae3D.py:44,208,211;
// Unet3D.py:22): the three largest kernels of the CAE training step when they run on the FFMA pipe.
//
//   dW[co][ci][kd][kh][kw] = sum_{n,od,oh,ow} dZ[n,od,oh,ow,co] * X'[n, od-pd+kd, oh-ph+kh, ow-pw+kw, ci]
//
// GEMM view: the voxels are the CONTRACTION index, so both operands are MN-major (channels contiguous, the 8 x 16-byte
// core matrix holds 8 consecutive voxels of 8 channels):
//   A (M side) = dZ tile, rows (term_y, co): the three bf16 terms of every fp32 gradient value are stacked along M (the
//                MMA computes all three for free: its cost is max(M,128) * N / 256 cycles), term t in TMEM lane quarter t;
//   B (N side) = X' tile (BatchNorm applied, zero padding written as zeros), columns (kh, ci) = N 48: the three input
//                rows oh+kh of one output row lie RS bytes apart per (row, channel half) — a uniform N-group stride —
//                while the kd and kw taps are start-address offsets of the descriptor (plane / 16-byte voxel shift);
//   K          = 16 consecutive output voxels of one row; the three bf16 terms of X' are three MMAs into the same
//                accumulator, so every product y_i * x_j (i, j <= 3) is formed: fp32-grade like the forward tier.
// One accumulator block of 48 TMEM columns per (kd, kw) = 432 columns, one issuing warp per block (MMAs of one thread
// retire one after the other, MMAs of different warps overlap).  The tensor core's accumulator
// truncates on every accumulation, so the blocks are drained every `drain_every` tiles (24 accumulations per tile) and
// summed in fp32 round-to-nearest in shared memory; the three y terms are folded smallest first at the end.
//
// Persistent CTA per SM, 21 warps: 0-3 drain TMEM (warp t = term t), 4-12 issue the MMAs of block (kd, kw), 13-20
// stage step i+1 (global -> BN -> exact 3-term split -> shared memory) while step i is multiplied.  A CTA walks a
// COLUMN (n, 4 rows, 32 columns) along the depth axis: the X' planes live in a ring of six slots, so every step stages
// one new input plane (6 x 34 voxels) and one dZ tile (4 x 32 voxels, double buffered) instead of three planes — the kd
// tap is the slot offset in the descriptor's start address.
// Partials of every CTA go to ws[cta][co][ci][27] (torch layout) and are folded in fixed order by wgrad_reduce_kernel.
#pragma once
#include "sp_conv_tc2.cuh"

namespace sp_wtc {

using namespace sp_tc;
using sp_tc2::mbar_arrive;
using sp_tc2::split8_trunc3;
using sp_tc2::tmem_ld48;

constexpr int TWW = 32, THW = 4;                   // output tile (w, h): 128 voxels = 8 K steps of 16
constexpr int XW = TWW + 2, XH = THW + 2;          // input halo tile
constexpr int RS = XW * 16;                        // bytes per (row, channel half) of the X tile = N-group stride
constexpr int X_PLANE_B = XH * 2 * RS;             // one input depth plane of one term
constexpr int NSLOT = 6;                           // ring of input planes (three in use, up to three being staged)
constexpr int X_TERM_B = NSLOT * X_PLANE_B;
constexpr int X_REGION_B = 3 * X_TERM_B;           // 117504
constexpr int PS = TWW * THW * 16;                 // bytes per (term, buffer, half) plane of the dZ tile = M-group stride
constexpr int A_REGION_B = 12 * PS;                // [term][buffer][half]: term t of a buffer starts 4 groups after term t-1
constexpr int NBLK = 9, BCOLS = 48;                // accumulator blocks (kd, kw) x columns (kh, ci)
constexpr int ACC_LD = 436;                        // floats per accumulator row (4 * odd: conflict-free 128-bit rows)
constexpr int ACC_B = 48 * ACC_LD * 4;
constexpr int W_EPI = 4, W_MMA = 9, W_STG = 8;
constexpr int NSTG = W_STG * 32;
constexpr int NTHREADS_W = (W_EPI + W_MMA + W_STG) * 32;      // 672
constexpr int XP_ITEMS = XH * XW * 2, NZ_ITEMS = TWW * THW * 2;   // items (8 channels of one voxel) per X plane / dZ tile
constexpr int PER_R = 3;
constexpr int N_BARS = 2 + 2 + NBLK + NBLK;
constexpr size_t SMEM_W = (size_t)A_REGION_B + X_REGION_B + ACC_B + N_BARS * 8 + 16;
static_assert(SMEM_W <= 227 * 1024, "wgrad tc: shared memory");
static_assert(XP_ITEMS % 2 == 0 && NSTG % 2 == 0, "a staging thread keeps one channel half");
static_assert(A_REGION_B + X_REGION_B >= 18 * PS, "the 16 M groups read from buffer 1 stay inside the A + X regions");

// kind::f16 instruction descriptor, D = f32, A = B = bf16, M = 128, both operands MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t idesc_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of the (term, buffer, half) plane of the dZ tile.  M = 128: [term][buffer][half], so term t starts at M
// group 4t = TMEM lane 32t; M = 64: [buffer][term][half], rows 16t.. of the 64-row accumulator = lanes 32t.. as well.
template <int MROWS>
__host__ __device__ constexpr int a_off(int term, int buf, int half) {
    return MROWS == 128 ? (term * 4 + buf * 2 + half) * PS : (buf * 6 + term * 2 + half) * PS;
}

template <int MROWS>
__global__ void __launch_bounds__(NTHREADS_W, 1)
wgrad3_tc_kernel(SpConvDesc d, int nPerG, int tiles_w, int tiles_h, int total_cols, int drain_every, int isstride, int osstride,
                 const float* __restrict__ X, const float* __restrict__ i_scale, const float* __restrict__ i_shift,
                 const float* __restrict__ dZ, const float* __restrict__ o_scale, const float* __restrict__ o_shift,
                 float* __restrict__ ws, long long* __restrict__ prof, int nt) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* a_reg = smem_raw;                                          // dZ terms, both buffers
    unsigned char* x_reg = smem_raw + A_REGION_B;                             // [term][slot][row][half][w] x 16 B
    float* acc = reinterpret_cast<float*>(smem_raw + A_REGION_B + X_REGION_B);    // [term][co][ACC_LD]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + A_REGION_B + X_REGION_B + ACC_B);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool pr = (prof != nullptr) && (blockIdx.x == 0);
    long long pw0 = 0, pw1 = 0, pwk = 0;

    for (int i = tid; i < 48 * ACC_LD; i += NTHREADS_W) acc[i] = 0.f;
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), NSTG); mbar_init(smem_u32(&bars[1]), NSTG);        // a_full
        mbar_init(smem_u32(&bars[2]), NBLK); mbar_init(smem_u32(&bars[3]), NBLK);        // a_empty: one commit per issuer
        for (int b = 0; b < NBLK; ++b) {
            mbar_init(smem_u32(&bars[4 + b]), 1);                                         // t_full[b]: issuer b
            mbar_init(smem_u32(&bars[4 + NBLK + b]), 3);                                  // t_empty[b]: drain warps 0..2
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
    const uint32_t t_full = smem_u32(&bars[4]), t_empty = smem_u32(&bars[4 + NBLK]);

    const int ncols = (total_cols - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // columns of this CTA
    const int nsteps = ncols * d.Do;
    const int ndrains = (nsteps + drain_every - 1) / drain_every;

    if (warp >= W_EPI + W_MMA) {
        // =================================================================== staging warps
        const int st = tid - (W_EPI + W_MMA) * 32;
        const int half = st & 1;
        const bool vec_i = (d.ldi % 4 == 0), vec_o = (d.ldo % 4 == 0);
        int it = 0, pc = 0;                                        // step and staged-plane counters of this CTA
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
            float bsc[8], bsh[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool okc = i_scale && (half * 8 + j < d.Ci);
                bsc[j] = okc ? i_scale[(int64_t)g * isstride + half * 8 + j] : 1.f;      // isstride: channels of the whole layer
                bsh[j] = okc ? i_shift[(int64_t)g * isstride + half * 8 + j] : 0.f;      // when X is a 16-channel slice of it
            }
            for (int od = 0; od < d.Do; ++od, ++it) {
                const int buf = it & 1, use = it >> 1;
                long long c0 = pr ? clock64() : 0;
                mbar_wait(a_empty + 8 * buf, (use & 1) ^ 1);       // the MMAs of step it-2 are done
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                // planes staged by this step: all three at the top of a column, afterwards the new deepest one
                const int np = (od == 0) ? 3 : 1;
                const int gd0 = od - d.pd + (3 - np);
                const int nx_items = np * XP_ITEMS, n_items = nx_items + NZ_ITEMS;
#pragma unroll 1
                for (int base = 0; base < n_items; base += NSTG * PER_R) {
                    float4 ra[PER_R], rb[PER_R];
                    int dsto[PER_R];                               // byte offset of term 0 in shared memory (-1: no item)
                    int flags[PER_R];                              // bit 0: inside the volume, bit 1: dZ item
#pragma unroll
                    for (int u = 0; u < PER_R; ++u) {
                        const int item = base + st + u * NSTG;
                        ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        rb[u] = ra[u];
                        dsto[u] = -1;
                        flags[u] = 0;
                        if (item < nx_items) {
                            int r = item >> 1;
                            const int wx = r % XW; r /= XW;
                            const int hy = r % XH;
                            const int p = r / XH;
                            const int slot = (pc + p) % NSLOT;
                            dsto[u] = ((slot * XH + hy) * 2 + half) * RS + wx * 16;
                            const int gd = gd0 + p, gh = ih0 + hy, gw = iw0 + wx;
                            const int c = half * 8;
                            if (gd >= 0 && gd < d.Di && gh >= 0 && gh < d.Hi && gw >= 0 && gw < d.Wi && c < d.Ci) {
                                flags[u] = 1;
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
                            const int v = (item - nx_items) >> 1;       // voxel of the output tile, row-major
                            dsto[u] = a_off<MROWS>(0, buf, half) + v * 16;
                            flags[u] = 2;
                            const int gh = oh0 + v / TWW, gw = ow0 + v % TWW;
                            const int c = half * 8;
                            if (gh < d.Ho && gw < d.Wo && c < d.Co) {
                                flags[u] = 3;
                                const float* pp = zn + (((int64_t)od * d.Ho + gh) * d.Wo + gw) * d.ldo + c;
                                if (vec_o && c + 8 <= d.Co) {
                                    ra[u] = *reinterpret_cast<const float4*>(pp);
                                    rb[u] = *reinterpret_cast<const float4*>(pp + 4);
                                } else {
                                    float e[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) e[j] = (c + j < d.Co) ? pp[j] : 0.f;
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
                        const bool isz = (flags[u] & 2) != 0;
                        if (flags[u] & 1) {
                            if (!isz) {
                                if (i_scale) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], bsc[j], bsh[j]);
                                }
                            } else if (o_scale) {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (half * 8 + j < d.Co)
                                        v[j] = fmaf(v[j], o_scale[(int64_t)g * osstride + half * 8 + j], o_shift[(int64_t)g * osstride + half * 8 + j]);
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
                        const int tstride = isz ? (a_off<MROWS>(1, 0, 0) - a_off<MROWS>(0, 0, 0)) : X_TERM_B;
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
        // =================================================================== MMA issue: warp 4 + b owns block b = (kd, kw)
        if (lane == 0) {
            const int b = warp - W_EPI, kd = b / 3, kw = b % 3;
            const uint32_t a_base = smem_u32(a_reg), x_base = smem_u32(x_reg);
            const uint32_t dcol = tmem_base + (uint32_t)(b * BCOLS);
            constexpr uint32_t IDESC = idesc_mn(MROWS, BCOLS);
            bool fresh = true;
            int drains = 0, pc = 0;
            for (int it = 0; it < nsteps; ++it) {
                const int buf = it & 1, use = it >> 1;
                pc += (it % d.Do == 0) ? 3 : 1;                   // planes staged up to and including this step
                long long c0 = pr ? clock64() : 0;
                mbar_wait(a_full + 8 * buf, use & 1);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                if (fresh && drains > 0) mbar_wait(t_empty + 8 * b, (drains - 1) & 1);
                long long c2 = pr ? clock64() : 0;
                pw1 += c2 - c1;
                tc_fence_after();
                // descriptors differ only in their start-address field (16-byte units)
                const int slot = (pc - 3 + kd) % NSLOT;           // plane od - pd + kd of this step
                const uint64_t da0 = umma_desc(a_base + (uint32_t)a_off<MROWS>(0, buf, 0), 128, PS);
                const uint64_t db0 = umma_desc(x_base + (uint32_t)(slot * X_PLANE_B + kw * 16), 128, RS);
#pragma unroll 1
                for (int r = 0; r < THW; ++r) {
#pragma unroll
                    for (int ks = 0; ks < TWW / 16; ++ks) {
                        const uint64_t da = da0 + (uint64_t)(r * TWW + ks * 16);
                        const uint64_t db = db0 + (uint64_t)((r * 2 * RS + ks * 256) >> 4);
                        for (int tx = nt - 1; tx >= 0; --tx) {              // smallest x term first
                            umma_bf16(dcol, da, db + (uint64_t)((tx * X_TERM_B) >> 4), IDESC, fresh ? 0u : 1u);
                            fresh = false;
                        }
                    }
                }
                umma_commit(a_empty + 8 * buf);
                if ((it + 1) % drain_every == 0 || it == nsteps - 1) {
                    umma_commit(t_full + 8 * b);
                    fresh = true;
                    ++drains;
                }
                if (pr) pwk += clock64() - c2;
            }
            if (pr && b == 0) { prof[0] = pw0; prof[1] = pw1; prof[2] = pwk; prof[3] = nsteps; }
        }
    } else if (warp < 3) {
        // =================================================================== drain: warp t holds the rows of y term t
        float* arow = acc + (size_t)(warp * 16 + (lane & 15)) * ACC_LD;
        for (int dr = 0; dr < ndrains; ++dr) {
#pragma unroll 1
            for (int b = 0; b < NBLK; ++b) {
                long long c0 = pr ? clock64() : 0;
                mbar_wait(t_full + 8 * b, dr & 1);
                long long c1 = pr ? clock64() : 0;
                pw0 += c1 - c0;
                tc_fence_after();
                float v[BCOLS];
                tmem_ld48(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * BCOLS), v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_empty + 8 * b);
                if (lane < 16) {
                    float4* ap = reinterpret_cast<float4*>(arow + b * BCOLS);
#pragma unroll
                    for (int j4 = 0; j4 < BCOLS / 4; ++j4) {
                        float4 a = ap[j4];
                        a.x += v[4 * j4]; a.y += v[4 * j4 + 1]; a.z += v[4 * j4 + 2]; a.w += v[4 * j4 + 3];
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
    // fold the three y terms (smallest first) and write this CTA's partial in torch layout dW[co][ci][tap]
    const int wn = d.Co * d.Ci * 27;
    float* wsp = ws + (int64_t)blockIdx.x * wn;
    for (int i = tid; i < wn; i += NTHREADS_W) {
        const int tap = i % 27, ci = (i / 27) % d.Ci, co = i / (27 * d.Ci);
        const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
        const int col = (kd * 3 + kw) * BCOLS + kh * 16 + ci;
        wsp[i] = (acc[(32 + co) * ACC_LD + col] + acc[(16 + co) * ACC_LD + col]) + acc[co * ACC_LD + col];
    }
}

struct WtcPlan {
    int tiles_w, tiles_h, grid;
    int64_t total;      // columns (n, 4 rows, 32 columns); each is walked along the depth axis
};
static inline WtcPlan plan(const SpConvDesc* d) {
    WtcPlan p;
    p.tiles_w = (d->Wo + TWW - 1) / TWW;
    p.tiles_h = (d->Ho + THW - 1) / THW;
    p.total = (int64_t)p.tiles_w * p.tiles_h * d->N;
    p.grid = sp_num_sms();
    if (p.grid > p.total) p.grid = (int)p.total;
    return p;
}

}  // namespace sp_wtc

#ifndef SP_WTC_MROWS
#define SP_WTC_MROWS 64
#endif

static inline bool sp_tc_wgrad_disabled() {
    static int v = -1;   // SP_DISABLE_TC_WGRAD=1 keeps the weight gradients on the FFMA tier
    if (v < 0) {
        const char* e = getenv("SP_DISABLE_TC_WGRAD");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static inline bool sp_tc_wgrad_supported(const SpConvDesc* d) {
    if (d->k != 3 || d->s != 1 || sp_tc_terms() == 0 || sp_tc_wgrad_disabled()) return false;
    if (d->Ci <= 8 || d->Ci > 16 || d->Co <= 8 || d->Co > 16) return false;
    if (d->pd > 2 || d->ph > 2 || d->pw > 2) return false;
    const sp_wtc::WtcPlan p = sp_wtc::plan(d);
    return p.total >= 16 && p.total < (1LL << 31) && d->Wo >= 24 && d->Do >= 8;
}

static inline size_t sp_tc_wgrad_workspace_bytes(const SpConvDesc* d) {
    if (!sp_tc_wgrad_supported(d)) return 0;
    return (size_t)sp_wtc::plan(d).grid * d->Co * d->Ci * 27 * sizeof(float);
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int chunks, int64_t wn, float* __restrict__ dw, float beta);

static inline int sp_tc_wgrad_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale, const float* i_shift,
                                     const float* oside, const float* o_scale, const float* o_shift, float* dw, float beta, float* ws,
                                     cudaStream_t st, long long* prof = nullptr, int drain_every = 2, int mrows = SP_WTC_MROWS) {
    using namespace sp_wtc;
    const WtcPlan p = plan(d);
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(wgrad3_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_W));
        SP_CUDA(cudaFuncSetAttribute(wgrad3_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_W));
        attr = true;
    }
    if (mrows == 64)
        wgrad3_tc_kernel<64><<<p.grid, NTHREADS_W, SMEM_W, st>>>(*d, nPerG, p.tiles_w, p.tiles_h, (int)p.total, drain_every, d->Ci, d->Co, iside,
                                                                 i_scale, i_shift, oside, o_scale, o_shift, ws, prof, sp_tc_terms() == 1 ? 1 : 3);
    else
        wgrad3_tc_kernel<128><<<p.grid, NTHREADS_W, SMEM_W, st>>>(*d, nPerG, p.tiles_w, p.tiles_h, (int)p.total, drain_every, d->Ci, d->Co, iside,
                                                                  i_scale, i_shift, oside, o_scale, o_shift, ws, prof, sp_tc_terms() == 1 ? 1 : 3);
    SP_LAUNCH_OK("wgrad3_tc_kernel");
    const int64_t wn = (int64_t)d->Co * d->Ci * 27;
    int64_t rb = (wn + 255) / 256;
    wgrad_reduce_kernel<<<(int)rb, 256, 0, st>>>(ws, p.grid, wn, dw, beta);
    SP_LAUNCH_OK("wgrad_reduce_kernel");
    return 0;
}

// ---- wider layers (Unet3D.py:19,22: 48 -> 16, 32 -> 32, 96 -> 32, ...): both sides run as slices of 16 channels through the
// kernel above (X + 16 c / dZ + 16 c' with the layer's row strides and its slices of the affine coefficients); every slice
// pair has its own per-CTA partials, which a scatter-reduce folds into dW[16 c' + co][16 c + ci][tap].
__global__ void wgrad_reduce_slice_kernel(const float* __restrict__ ws, int chunks, int cos, int cs, int Ci, int co0, int ci0,
                                          float* __restrict__ dw, float beta) {
    const int wn = cos * cs * 27;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wn) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += ws[(int64_t)c * wn + i];
    const int tap = i % 27, ci = (i / 27) % cs, co = i / (27 * cs);
    float* o = dw + ((int64_t)(co0 + co) * Ci + ci0 + ci) * 27 + tap;
    *o = (beta != 0.f ? beta * *o : 0.f) + s;
}

static inline bool sp_tc_wgrad_sliced_supported(const SpConvDesc* d) {
    if (d->k != 3 || d->s != 1 || sp_tc_terms() == 0 || sp_tc_wgrad_disabled()) return false;
    if (d->Ci <= 24 && d->Co <= 24) return false;                    // the single-launch kernels take these
    // Measured on the U-Net (B200): slicing pays for 48 -> 16 on 30x130x130 (2.48 -> 2.02 ms) but not for the spatially small
    // wide layers, where every slice pair re-stages both tiles and pays the pipeline fill of 148 persistent CTAs
    // (96 -> 32 on 18x68x68: 1.86 -> 2.46 ms, 32 -> 32 on 28x78x78: 0.97 -> 1.17 ms): at most three slice pairs are taken.
    if (((d->Ci + 15) / 16) * ((d->Co + 15) / 16) > 3) return false;
    if (d->Ci <= 8 || d->Ci > 96 || d->Co <= 8 || d->Co > 64 || d->ldi % 4 != 0 || d->ldo % 4 != 0) return false;
    if (d->pd > 2 || d->ph > 2 || d->pw > 2) return false;
    const sp_wtc::WtcPlan p = sp_wtc::plan(d);
    return p.total >= 16 && p.total < (1LL << 31) && d->Wo >= 24 && d->Do >= 8;
}

static inline size_t sp_tc_wgrad_sliced_workspace_bytes(const SpConvDesc* d) {
    if (!sp_tc_wgrad_sliced_supported(d)) return 0;
    return (size_t)sp_wtc::plan(d).grid * 16 * 16 * 27 * sizeof(float);       // one slice pair at a time (stream-ordered reuse)
}

static inline int sp_tc_wgrad_sliced_launch(const SpConvDesc* d, int nPerG, const float* iside, const float* i_scale,
                                            const float* i_shift, const float* oside, const float* o_scale, const float* o_shift,
                                            float* dw, float beta, float* ws, cudaStream_t st, int drain_every = 2) {
    using namespace sp_wtc;
    const WtcPlan p = plan(d);
    static bool attr = false;
    if (!attr) {
        SP_CUDA(cudaFuncSetAttribute(wgrad3_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_W));
        attr = true;
    }
    const int nsi = (d->Ci + 15) / 16, nso = (d->Co + 15) / 16;
    for (int co = 0; co < nso; ++co)
        for (int c = 0; c < nsi; ++c) {
            SpConvDesc s = *d;
            s.Ci = (d->Ci - 16 * c < 16) ? d->Ci - 16 * c : 16;
            s.Co = (d->Co - 16 * co < 16) ? d->Co - 16 * co : 16;
            wgrad3_tc_kernel<64><<<p.grid, NTHREADS_W, SMEM_W, st>>>(s, nPerG, p.tiles_w, p.tiles_h, (int)p.total, drain_every, d->Ci, d->Co,
                                                                     iside + 16 * c, i_scale ? i_scale + 16 * c : nullptr,
                                                                     i_shift ? i_shift + 16 * c : nullptr, oside + 16 * co,
                                                                     o_scale ? o_scale + 16 * co : nullptr,
                                                                     o_shift ? o_shift + 16 * co : nullptr, ws, nullptr, sp_tc_terms() == 1 ? 1 : 3);
            SP_LAUNCH_OK("wgrad3_tc_kernel");
            const int wn = s.Co * s.Ci * 27;
            wgrad_reduce_slice_kernel<<<(wn + 255) / 256, 256, 0, st>>>(ws, p.grid, s.Co, s.Ci, d->Ci, 16 * co, 16 * c, dw, beta);
            SP_LAUNCH_OK("wgrad_reduce_slice_kernel");
        }
    return 0;
}
