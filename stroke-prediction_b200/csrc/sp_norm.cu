// sp_norm.cu — BatchNorm3d statistics, finalisation and backward (replaces nn.BatchNorm3d at
// common/model/Cae3D.py:40..217 and common/model/Unet3D.py:18,21; semantics pinned in SURVEY App. D).
//
// The normalised tensor is never written: forward produces per-(group,channel) scale/shift that the consuming
// convolution applies while staging its input; backward produces the coefficients of
//   gx = A*((gxh - m1) - (x - mu)*k)   (gxh = gradient w.r.t. the BN output)
// which sp_bn_act_bwd_apply fuses with the derivative of the activation that produced x.
//
// All channel reductions accumulate in fp64 (the variance sum x^2 - n*mean^2 cancels catastrophically in fp32 for
// the raw CBV/TTD inputs whose mean is far from 0), first per thread, then per CTA through shared memory, then
// with one fp64 atomicAdd per (CTA, channel).
#include "sp_common.cuh"

namespace {

// Layout of the reduction: x is [G][rows][ld] with rows = (N/G)*vox.  A CTA owns a contiguous slab of rows of one
// group.  Thread t owns column c = t % lanesC and row-lane r = t / lanesC (coalesced along the channel axis; rows
// are contiguous when ld == C).  TWO = 1: sums of x and x*x; TWO = 2: sums of a and a*b (BN backward).
template <int MODE>
__global__ void __launch_bounds__(256)
channel_sums_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, int64_t rows_per_group,
                    int C, int64_t rows_per_block, int blocks_per_group, double* __restrict__ sums) {
    __shared__ double sm0[256];
    __shared__ double sm1[256];
    const int g = blockIdx.x / blocks_per_group;
    const int bg = blockIdx.x % blocks_per_group;
    const int lanesC = C < 256 ? C : 256;
    const int R = 256 / lanesC;
    const int c_l = threadIdx.x % lanesC, r_l = threadIdx.x / lanesC;
    const int64_t r0 = (int64_t)bg * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < rows_per_group) ? r0 + rows_per_block : rows_per_group;
    const float* ag = a + (int64_t)g * rows_per_group * lda;
    const float* bgp = (MODE == 2) ? b + (int64_t)g * rows_per_group * ldb : nullptr;
    for (int cb = 0; cb < C; cb += lanesC) {   // uniform trip count: every thread reaches the barriers below
        const int c = cb + c_l;
        const bool live = (c < C) && (r_l < R);
        double s0 = 0.0, s1 = 0.0;
        if (live) {
            // straight fp64 accumulation: the product of two floats is exact in double, so sum(a*b) - mu*sum(a) and
            // sum(x^2) - n*mean^2 cancel without loss whatever |mean| / std of the tensor is (B200 runs DFMA at half the
            // FFMA rate; the kernel stays HBM-bound).  fp32 partial sums here cost |mean|/std digits of the BN backward
            // dot product and showed up as 4-5x the CPU's gradient noise on nets with non-trivial BN affine parameters.
            for (int64_t r = r0 + r_l; r < r1; r += R) {
                const double va = (double)ag[r * lda + c];
                const double vb = (MODE == 1) ? va : (double)bgp[r * ldb + c];
                s0 += va;
                s1 = fma(va, vb, s1);
            }
        }
        sm0[threadIdx.x] = s0;
        sm1[threadIdx.x] = s1;
        __syncthreads();
        if (r_l == 0 && c < C) {
            for (int q = 1; q < R; ++q) {
                s0 += sm0[q * lanesC + c_l];
                s1 += sm1[q * lanesC + c_l];
            }
            atomicAdd(&sums[((int64_t)g * C + c) * 2 + 0], s0);
            atomicAdd(&sums[((int64_t)g * C + c) * 2 + 1], s1);
        }
        __syncthreads();
    }
}

// float4 variant (C, lda, ldb multiples of 4): thread = (channel quad, row lane); four rows in flight per thread so that
// enough bytes are outstanding per SM to reach the HBM roof (the scalar kernel above keeps ~8 KB/SM in flight).
template <int MODE>
__global__ void __launch_bounds__(256)
channel_sums_vec_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, int64_t rows_per_group,
                        int C4, int64_t rows_per_block, int blocks_per_group, double* __restrict__ sums) {
    __shared__ double sm[256][8];
    const int g = blockIdx.x / blocks_per_group;
    const int bg = blockIdx.x % blocks_per_group;
    const int lanesQ = C4 < 256 ? C4 : 256;
    const int R = 256 / lanesQ;
    const int c_l = threadIdx.x % lanesQ, r_l = threadIdx.x / lanesQ;
    const int64_t r0 = (int64_t)bg * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < rows_per_group) ? r0 + rows_per_block : rows_per_group;
    const float* ag = a + (int64_t)g * rows_per_group * lda;
    const float* bgp = (MODE == 2) ? b + (int64_t)g * rows_per_group * ldb : nullptr;
    for (int cb = 0; cb < C4; cb += lanesQ) {   // uniform trip count: every thread reaches the barriers below
        const int q = cb + c_l;
        const bool live = (q < C4) && (r_l < R);
        double s0[4] = {0.0, 0.0, 0.0, 0.0}, s1[4] = {0.0, 0.0, 0.0, 0.0};
        if (live) {
            auto add = [&](const float4& va, const float4& vb) {
                const double a0 = va.x, a1 = va.y, a2 = va.z, a3 = va.w;
                s0[0] += a0; s0[1] += a1; s0[2] += a2; s0[3] += a3;
                s1[0] = fma(a0, (double)vb.x, s1[0]); s1[1] = fma(a1, (double)vb.y, s1[1]);
                s1[2] = fma(a2, (double)vb.z, s1[2]); s1[3] = fma(a3, (double)vb.w, s1[3]);
            };
            int64_t r = r0 + r_l;
            constexpr int U = (MODE == 1) ? 8 : 4;          // eight 128-bit loads in flight per thread in both modes
            for (; r + (U - 1) * (int64_t)R < r1; r += U * (int64_t)R) {
                float4 va[U], vb[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    va[u] = sp_ldg_stream(reinterpret_cast<const float4*>(ag + (r + (int64_t)u * R) * lda + q * 4));
                    vb[u] = (MODE == 1) ? va[u] : sp_ldg_stream(reinterpret_cast<const float4*>(bgp + (r + (int64_t)u * R) * ldb + q * 4));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) add(va[u], vb[u]);
            }
            for (; r < r1; r += R) {
                const float4 va = *reinterpret_cast<const float4*>(ag + r * lda + q * 4);
                const float4 vb = (MODE == 1) ? va : *reinterpret_cast<const float4*>(bgp + r * ldb + q * 4);
                add(va, vb);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sm[threadIdx.x][j] = s0[j];
            sm[threadIdx.x][4 + j] = s1[j];
        }
        __syncthreads();
        if (r_l == 0 && q < C4) {
            for (int k = 1; k < R; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s0[j] += sm[k * lanesQ + c_l][j];
                    s1[j] += sm[k * lanesQ + c_l][4 + j];
                }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&sums[((int64_t)g * C4 * 4 + q * 4 + j) * 2 + 0], s0[j]);
                atomicAdd(&sums[((int64_t)g * C4 * 4 + q * 4 + j) * 2 + 1], s1[j]);
            }
        }
        __syncthreads();
    }
}

// One thread per channel; groups are visited in order so the running statistics see G sequential momentum updates
// exactly like G separate nn.BatchNorm3d calls (Cae3D.py:105-108: core, penu, lesion[, interpolation]).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C, int G, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int64_t* __restrict__ nbt, float momentum, float eps,
                                   int training, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && training && nbt) *nbt += G;
    if (c >= C) return;
    const float gm = gamma ? gamma[c] : 1.f;
    const float bt = beta ? beta[c] : 0.f;
    float rm = running_mean ? running_mean[c] : 0.f;
    float rv = running_var ? running_var[c] : 1.f;
    for (int g = 0; g < G; ++g) {
        float mean, invstd;
        if (training) {
            const double s0 = sums[((int64_t)g * C + c) * 2 + 0];
            const double s1 = sums[((int64_t)g * C + c) * 2 + 1];
            const double m = s0 / count;
            double var = s1 / count - m * m;
            if (var < 0.0) var = 0.0;
            mean = (float)m;
            invstd = (float)(1.0 / sqrt(var + (double)eps));
            const float unbiased = (float)(count > 1.0 ? var * (count / (count - 1.0)) : var);
            rm = (1.f - momentum) * rm + momentum * mean;
            rv = (1.f - momentum) * rv + momentum * unbiased;
        } else {
            mean = rm;
            invstd = 1.f / sqrtf(rv + eps);
        }
        const float sc = gm * invstd;
        scale[g * C + c] = sc;
        // shift is formed in double from the ROUNDED scale and mean, so that x*sc + shift == (x - mean)*sc + beta up to
        // the one rounding of shift itself
        shift[g * C + c] = (float)((double)bt - (double)mean * (double)sc);
        mean_out[g * C + c] = mean;
        invstd_out[g * C + c] = invstd;
    }
    if (training) {
        if (running_mean) running_mean[c] = rm;
        if (running_var) running_var[c] = rv;
    }
}

// coef layout [4][G][C]: A, m1, mu, k  with   gx = A * ((gxh - m1) - (x - mu) * k)
//   training: A = gamma*invstd, m1 = mean(gxh), mu = batch mean, k = invstd^2 * mean(gxh * (x - mu))
//   eval    : A = gamma*invstd(running), m1 = 0, k = 0
// The subtraction (gxh - m1) is done first and in fp32 exactly like ATen's batch_norm_backward: when the incoming
// gradient has a large common mode (the hinge term seeds -2/N on half of all voxels) it cancels without rounding;
// folding m1 into a pre-rounded constant would lose |m1| / |gxh - m1| digits.
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ bsums, double count, int C, int G,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, int training, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float beta_acc, float* __restrict__ coef) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float gm = gamma ? gamma[c] : 1.f;
    double dg = 0.0, db = 0.0;
    for (int g = 0; g < G; ++g) {
        const double s1 = bsums[((int64_t)g * C + c) * 2 + 0];       // sum gxh
        const double sx = bsums[((int64_t)g * C + c) * 2 + 1];       // sum gxh*x
        const double mu = (double)mean[g * C + c];
        const double is = (double)invstd[g * C + c];
        const double dotp = sx - mu * s1;                            // sum gxh*(x - mu)
        dg += is * dotp;                                             // sum gxh*xhat
        db += s1;
        coef[(0 * G + g) * C + c] = (float)((double)gm * is);
        coef[(1 * G + g) * C + c] = training ? (float)(s1 / count) : 0.f;
        coef[(2 * G + g) * C + c] = (float)mu;
        coef[(3 * G + g) * C + c] = training ? (float)(is * is * dotp / count) : 0.f;
    }
    if (dgamma) dgamma[c] = (beta_acc == 0.f) ? (float)dg : fmaf(beta_acc, dgamma[c], (float)dg);
    if (dbeta) dbeta[c] = (beta_acc == 0.f) ? (float)db : fmaf(beta_acc, dbeta[c], (float)db);
}

// out[v][c] (+)= A * ((gxh - m1) - (x - mu) * k) * act'(x); elementwise over [N*vox][C]
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const float* __restrict__ gxh, int ldg, const float* __restrict__ x, int ldx,
                        const float* __restrict__ coef, int64_t rows, int64_t rows_per_group, int C, int G, int act,
                        float alpha, float* __restrict__ out, int ldout, int accumulate) {
    const int64_t total = rows * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t r = i / C;
        const float gv = gxh[r * ldg + c];
        float res = gv;
        float xv = 0.f;
        if (coef || act != SP_ACT_NONE) xv = x[r * ldx + c];
        if (coef) {
            const int g = (int)(r / rows_per_group);
            const float A = coef[(0 * G + g) * C + c], m1 = coef[(1 * G + g) * C + c];
            const float mu = coef[(2 * G + g) * C + c], k = coef[(3 * G + g) * C + c];
            res = A * ((gv - m1) - (xv - mu) * k);
        }
        res *= sp_act_bwd(xv, act, alpha);
        float* o = out + r * ldout + c;
        *o = accumulate ? (*o + res) : res;
    }
}

// float4 variant of the apply kernel (C and all lds multiples of 4).  Thread = (channel quad, row lane), groups are walked
// one after the other so the four coefficients of the thread's channels sit in registers.  colsum != NULL: also
// accumulate the per-channel column sums of the values written (fp64; = bias gradient of the convolution whose output
// gradient this kernel produces, Learner.py:121 loss.backward()).
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_vec_kernel(const float* __restrict__ gxh, int ldg, const float* __restrict__ x, int ldx,
                            const float* __restrict__ coef, int64_t rows_per_group, int C4, int G, int act, float alpha,
                            float* __restrict__ out, int ldout, int accumulate, double* __restrict__ colsum) {
    __shared__ double sm[256][4];
    const int C = C4 * 4;
    const int lanesQ = C4 < 256 ? C4 : 256;
    const int R = 256 / lanesQ;
    const int c_l = threadIdx.x % lanesQ, r_l = threadIdx.x / lanesQ;
    const bool need_x = (coef != nullptr) || (act != SP_ACT_NONE);
    for (int cb = 0; cb < C4; cb += lanesQ) {
        const int q = cb + c_l;
        const bool live = (q < C4) && (r_l < R);
        double cs[4] = {0.0, 0.0, 0.0, 0.0};
        if (live) {
            for (int g = 0; g < G; ++g) {
                float4 A = make_float4(1.f, 1.f, 1.f, 1.f), m1 = make_float4(0.f, 0.f, 0.f, 0.f), mu = m1, k = m1;
                if (coef) {
                    A = *reinterpret_cast<const float4*>(coef + (0 * G + g) * C + q * 4);
                    m1 = *reinterpret_cast<const float4*>(coef + (1 * G + g) * C + q * 4);
                    mu = *reinterpret_cast<const float4*>(coef + (2 * G + g) * C + q * 4);
                    k = *reinterpret_cast<const float4*>(coef + (3 * G + g) * C + q * 4);
                }
                const int64_t rbeg = (int64_t)g * rows_per_group, rend = rbeg + rows_per_group;
                const int64_t step = (int64_t)gridDim.x * R;
                auto one = [&](int64_t r, const float4& gv, const float4& xv) {
                    float4 res = gv;
                    if (coef) {
                        res.x = A.x * ((gv.x - m1.x) - (xv.x - mu.x) * k.x);
                        res.y = A.y * ((gv.y - m1.y) - (xv.y - mu.y) * k.y);
                        res.z = A.z * ((gv.z - m1.z) - (xv.z - mu.z) * k.z);
                        res.w = A.w * ((gv.w - m1.w) - (xv.w - mu.w) * k.w);
                    }
                    res.x *= sp_act_bwd(xv.x, act, alpha); res.y *= sp_act_bwd(xv.y, act, alpha);
                    res.z *= sp_act_bwd(xv.z, act, alpha); res.w *= sp_act_bwd(xv.w, act, alpha);
                    float4* o = reinterpret_cast<float4*>(out + r * ldout + q * 4);
                    if (accumulate) {
                        const float4 prev = *o;
                        res.x += prev.x; res.y += prev.y; res.z += prev.z; res.w += prev.w;
                    }
                    *o = res;
                    cs[0] += (double)res.x; cs[1] += (double)res.y; cs[2] += (double)res.z; cs[3] += (double)res.w;
                };
                int64_t r = rbeg + (int64_t)blockIdx.x * R + r_l;
                for (; r + step < rend; r += 2 * step) {
                    const float4 g0 = sp_ldg_stream(reinterpret_cast<const float4*>(gxh + r * ldg + q * 4));
                    const float4 g1 = sp_ldg_stream(reinterpret_cast<const float4*>(gxh + (r + step) * ldg + q * 4));
                    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
                    if (need_x) {
                        x0 = sp_ldg_stream(reinterpret_cast<const float4*>(x + r * ldx + q * 4));
                        x1 = sp_ldg_stream(reinterpret_cast<const float4*>(x + (r + step) * ldx + q * 4));
                    }
                    one(r, g0, x0);
                    one(r + step, g1, x1);
                }
                for (; r < rend; r += step) {
                    const float4 g0 = sp_ldg_stream(reinterpret_cast<const float4*>(gxh + r * ldg + q * 4));
                    float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (need_x) x0 = sp_ldg_stream(reinterpret_cast<const float4*>(x + r * ldx + q * 4));
                    one(r, g0, x0);
                }
            }
        }
        if (colsum) {     // uniform branch
#pragma unroll
            for (int j = 0; j < 4; ++j) sm[threadIdx.x][j] = cs[j];
            __syncthreads();
            if (r_l == 0 && q < C4) {
                for (int kk = 1; kk < R; ++kk)
#pragma unroll
                    for (int j = 0; j < 4; ++j) cs[j] += sm[kk * lanesQ + c_l][j];
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(&colsum[q * 4 + j], cs[j]);
            }
            __syncthreads();
        }
    }
}

// BatchNorm parameter gradients of a network's FIRST unit without its dgrad.  The input of that unit is data, so the gradient with
// respect to the BN output (a full transposed correlation, 0.91 ms of the U-Net step, plus its reduction pass) is only needed for
//   dgamma[ci] = sum_u dXbn[ci,u] * xhat[ci,u],   dbeta[ci] = sum_u dXbn[ci,u].
// For a convolution without padding every tap of every output voxel reads a real input voxel, so with dXbn = corrT(dZ, W)
//   dgamma[ci] = sum_{co,tap} W[co,ci,tap] * dWhat[co,ci,tap],   dWhat = sum_v dZ[co,v] * xhat[ci,v+tap]   (the weight gradient taken
//                                                                against the NORMALISED input: scale = invstd, shift = -mean*invstd)
//   dbeta[ci]  = sum_{co,tap} W[co,ci,tap] * S[co],              S[co] = sum_v dZ[co,v]                     (the bias gradient, fp64)
// and the layer's own weight gradient follows from xbn = gamma * xhat + beta:  dW = gamma[ci] * dWhat + beta[ci] * S[co].
// With zero padding (applied AFTER BatchNorm, Cae3D.py:40-41) a tap of a border voxel may read padding: S becomes per tap,
// S[co,tap] = S[co] - E[co,tap] with E = the sum of dZ over the output voxels whose tap falls outside (border_tap_sums_kernel).
// The sums run over all statistics groups at once (gamma / beta are shared, mean / invstd enter through dWhat).  One block per ci.
__global__ void __launch_bounds__(128)
bn_grads_from_wgrad_kernel(const float* __restrict__ W, const float* __restrict__ dWhat, const double* __restrict__ colsum,
                           const double* __restrict__ tap_excl, const float* __restrict__ gamma, const float* __restrict__ beta, int Co, int Ci, int k3,
                           float* __restrict__ dW, float beta_dw, float* __restrict__ dgamma, float* __restrict__ dbeta, float beta_acc) {
    __shared__ double sm[2][4];
    const int ci = blockIdx.x;
    const double gm = (double)gamma[ci], bt = (double)beta[ci];
    double sg = 0.0, sb = 0.0;
    for (int i = threadIdx.x; i < Co * k3; i += blockDim.x) {
        const int co = i / k3, tap = i - co * k3;
        const int64_t idx = ((int64_t)co * Ci + ci) * k3 + tap;
        const double w = (double)W[idx], dh = (double)dWhat[idx];
        const double cs = colsum[co] - (tap_excl ? tap_excl[co * k3 + tap] : 0.0);      // S[co, tap]: the voxels whose tap reads a real input
        sg = fma(w, dh, sg);
        sb = fma(w, cs, sb);
        const float v = (float)(gm * dh + bt * cs);
        dW[idx] = (beta_dw == 0.f) ? v : fmaf(beta_dw, dW[idx], v);
    }
    sg = sp_warp_sum(sg);
    sb = sp_warp_sum(sb);
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = sg; sm[1][threadIdx.x >> 5] = sb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float dg = (float)(sm[0][0] + sm[0][1] + sm[0][2] + sm[0][3]);
        const float db = (float)(sm[1][0] + sm[1][1] + sm[1][2] + sm[1][3]);
        if (dgamma) dgamma[ci] = (beta_acc == 0.f) ? dg : fmaf(beta_acc, dgamma[ci], dg);
        if (dbeta) dbeta[ci] = (beta_acc == 0.f) ? db : fmaf(beta_acc, dbeta[ci], db);
    }
}

// E[co][tap] = sum of gz[co, v] over the output voxels v of a 3x3x3 stride-1 convolution whose tap (kd, kh, kw) reads zero padding
// (input coordinate o + k - p outside [0, I)).  Only border voxels contribute: a block walks (n, od, oh) rows and reads whole rows
// on the d / h borders, the first / last pw voxels otherwise.  Thread = (output channel, w lane); 27 fp64 accumulators per thread.
__global__ void __launch_bounds__(256)
border_tap_sums_kernel(const float* __restrict__ gz, int ldz, int N, int Do, int Ho, int Wo, int Co, int CoP2, int pd, int ph, int pw,
                       int Di, int Hi, int Wi, double* __restrict__ E) {
    __shared__ double sm[256];
    const int co = threadIdx.x % CoP2, wl = threadIdx.x / CoP2, lanes = 256 / CoP2;
    double acc[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) acc[t] = 0.0;
    // Candidate rows only (walking all N*Do*Ho rows cost 0.2 ms of index arithmetic on the CAE's first layer): [A] every row of the
    // nd planes on the d border, [B] the nh rows on the h border of the other planes, [C] (pw > 0) the remaining rows, of which only
    // the first / last pw voxels are read.
    const int nd = (2 * pd < Do) ? 2 * pd : Do, nh = (2 * ph < Ho) ? 2 * ph : Ho;
    const int perA = nd * Ho, perB = (Do - nd) * nh, perC = (pw > 0) ? (Do - nd) * (Ho - nh) : 0;
    const int totA = N * perA, totB = N * perB, totC = N * perC;
    for (int idx = blockIdx.x; idx < totA + totB + totC; idx += gridDim.x) {
        int n, od, oh;
        if (idx < totA) {
            n = idx / perA;
            const int rem = idx - n * perA, k = rem / Ho;
            oh = rem - k * Ho;
            od = (nd == Do || k < pd) ? k : Do - 2 * pd + k;
        } else if (idx < totA + totB) {
            const int i = idx - totA;
            n = i / perB;
            const int rem = i - n * perB, kd = rem / nh, j = rem - kd * nh;
            od = pd + kd;
            oh = (nh == Ho || j < ph) ? j : Ho - 2 * ph + j;
        } else {
            const int i = idx - totA - totB, perRow = Ho - nh;
            n = i / perC;
            const int rem = i - n * perC, kd = rem / perRow;
            od = pd + kd;
            oh = ph + (rem - kd * perRow);
        }
        const int64_t r = ((int64_t)n * Do + od) * Ho + oh;
        unsigned md = 0, mh = 0;                                  // bit k: tap k of this axis reads padding
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (od + k - pd < 0 || od + k - pd >= Di) md |= 1u << k;
            if (oh + k - ph < 0 || oh + k - ph >= Hi) mh |= 1u << k;
        }
        const bool whole = (md | mh) != 0u || 2 * pw >= Wo;
        if (!whole && pw == 0) continue;
        const float* rowp = gz + r * (int64_t)Wo * ldz;
        const int cnt = whole ? Wo : 2 * pw;
        constexpr int U = 8;                                      // loads in flight per thread (the kernel is pure load latency)
        for (int j0 = wl; j0 < cnt; j0 += U * lanes) {
            float v[U];
            unsigned m[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u * lanes;
                const int ow = whole ? j : (j < pw ? j : Wo - 2 * pw + j);
                unsigned mw = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (ow + k - pw < 0 || ow + k - pw >= Wi) mw |= 1u << k;
                const bool live = j < cnt && co < Co && (md | mh | mw) != 0u;
                m[u] = live ? mw : 0xffffffffu;
                v[u] = live ? __ldg(rowp + (int64_t)ow * ldz + co) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (m[u] == 0xffffffffu) continue;
                const double dv = (double)v[u];
#pragma unroll
                for (int t = 0; t < 27; ++t)
                    if (((md >> (t / 9)) | (mh >> ((t / 3) % 3)) | (m[u] >> (t % 3))) & 1u) acc[t] += dv;
            }
        }
    }
#pragma unroll                   // (static indices: acc stays in registers)
    for (int t = 0; t < 27; ++t) {
        sm[threadIdx.x] = acc[t];
        __syncthreads();
        if (wl == 0 && co < Co) {
            double sum = 0.0;
            for (int l = 0; l < lanes; ++l) sum += sm[l * CoP2 + co];
            // same-address fp64 atomics serialise (~0.2 us each): the blocks are spread over SP_TAP_EXCL_REPLICAS copies of E
            if (sum != 0.0) atomicAdd(&E[(size_t)(blockIdx.x % SP_TAP_EXCL_REPLICAS) * 27 * Co + co * 27 + t], sum);
        }
        __syncthreads();
    }
}

__global__ void fold_replicas_kernel(double* __restrict__ E, int n, int replicas) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = E[i];
    for (int k = 1; k < replicas; ++k) s += E[(size_t)k * n + i];
    E[i] = s;
}

__global__ void colsum_to_bias_kernel(const double* __restrict__ acc, int C, float* __restrict__ db, float beta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) db[c] = (beta == 0.f) ? (float)acc[c] : fmaf(beta, db[c], (float)acc[c]);
}

// column sums of a [rows][ld] matrix (scalar layout fallback of the fused bias gradient)
__global__ void __launch_bounds__(256)
channel_sums_colonly_kernel(const float* __restrict__ a, int lda, int64_t rows, int C, int64_t rows_per_block, double* __restrict__ colsum) {
    __shared__ double sm[256];
    const int lanesC = C < 256 ? C : 256;
    const int R = 256 / lanesC;
    const int c_l = threadIdx.x % lanesC, r_l = threadIdx.x / lanesC;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < rows) ? r0 + rows_per_block : rows;
    for (int cb = 0; cb < C; cb += lanesC) {
        const int c = cb + c_l;
        double s = 0.0;
        if (r_l < R && c < C)
            for (int64_t r = r0 + r_l; r < r1; r += R) s += (double)a[r * lda + c];
        sm[threadIdx.x] = s;
        __syncthreads();
        if (r_l == 0 && c < C) {
            for (int q = 1; q < R; ++q) s += sm[q * lanesC + c_l];
            atomicAdd(&colsum[c], s);
        }
        __syncthreads();
    }
}

int launch_sums(int mode, const float* a, int lda, const float* b, int ldb, int N, int64_t vox, int C, int G,
                double* sums, cudaStream_t st) {
    const int64_t rows_per_group = (int64_t)(N / G) * vox;
    SP_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * G * C, st));
    // ~8 CTAs per SM over all groups; at least 16 rows per CTA
    int64_t bpg = sp_cdiv((int64_t)sp_num_sms() * 8, G);
    const int64_t maxb = sp_cdiv(rows_per_group, 16);
    if (bpg > maxb) bpg = maxb;
    if (bpg < 1) bpg = 1;
    const int64_t rpb = sp_cdiv(rows_per_group, bpg);
    bpg = sp_cdiv(rows_per_group, rpb);
    const bool vec = (C % 4 == 0) && (lda % 4 == 0) && (mode == 1 || ldb % 4 == 0);
    if (vec && mode == 1)
        channel_sums_vec_kernel<1><<<(int)(bpg * G), 256, 0, st>>>(a, lda, nullptr, 0, rows_per_group, C / 4, rpb, (int)bpg, sums);
    else if (vec)
        channel_sums_vec_kernel<2><<<(int)(bpg * G), 256, 0, st>>>(a, lda, b, ldb, rows_per_group, C / 4, rpb, (int)bpg, sums);
    else if (mode == 1)
        channel_sums_kernel<1><<<(int)(bpg * G), 256, 0, st>>>(a, lda, nullptr, 0, rows_per_group, C, rpb, (int)bpg, sums);
    else
        channel_sums_kernel<2><<<(int)(bpg * G), 256, 0, st>>>(a, lda, b, ldb, rows_per_group, C, rpb, (int)bpg, sums);
    SP_LAUNCH_OK("channel_sums_kernel");
    return 0;
}

}  // namespace

extern "C" {

int sp_bn_stats(const float* x, int N, int64_t vox, int C, int ld, int G, double* sums, void* stream) {
    SP_REQUIRE(x && sums, "sp_bn_stats: NULL pointer");
    SP_REQUIRE(N > 0 && vox > 0 && C > 0 && ld >= C && G >= 1 && N % G == 0, "sp_bn_stats: bad shape N=%d C=%d ld=%d G=%d", N, C, ld, G);
    return launch_sums(1, x, ld, nullptr, 0, N, vox, C, G, sums, sp_stream(stream));
}

int sp_bn_finalize(const double* sums, int64_t count_per_group, int C, int G, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, int64_t* nbt, float momentum, float eps, int training,
                   float* scale, float* shift, float* mean, float* invstd, void* stream) {
    SP_REQUIRE(scale && shift && mean && invstd, "sp_bn_finalize: NULL output");
    SP_REQUIRE(!training || sums, "sp_bn_finalize: training mode needs sums");
    SP_REQUIRE(training || (running_mean && running_var), "sp_bn_finalize: eval mode needs running statistics");
    SP_REQUIRE(C > 0 && G >= 1 && count_per_group > 0, "sp_bn_finalize: bad shape");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, sp_stream(stream)>>>(sums, (double)count_per_group, C, G, gamma, beta,
                                                                    running_mean, running_var, nbt, momentum, eps,
                                                                    training, scale, shift, mean, invstd);
    SP_LAUNCH_OK("bn_finalize_kernel");
    return 0;
}

int sp_bn_bwd_reduce(const float* gxh, int ldg, const float* x, int ldx, int N, int64_t vox, int C, int G, double* bsums,
                     void* stream) {
    SP_REQUIRE(gxh && x && bsums, "sp_bn_bwd_reduce: NULL pointer");
    SP_REQUIRE(N > 0 && vox > 0 && C > 0 && ldg >= C && ldx >= C && G >= 1 && N % G == 0, "sp_bn_bwd_reduce: bad shape");
    return launch_sums(2, gxh, ldg, x, ldx, N, vox, C, G, bsums, sp_stream(stream));
}

int sp_bn_bwd_finalize(const double* bsums, int64_t count_per_group, int C, int G, const float* gamma, const float* mean,
                       const float* invstd, int training, float* dgamma, float* dbeta, float beta_acc, float* coef,
                       void* stream) {
    SP_REQUIRE(bsums && mean && invstd && coef, "sp_bn_bwd_finalize: NULL pointer");
    SP_REQUIRE(C > 0 && G >= 1 && count_per_group > 0, "sp_bn_bwd_finalize: bad shape");
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, sp_stream(stream)>>>(bsums, (double)count_per_group, C, G, gamma, mean,
                                                                        invstd, training, dgamma, dbeta, beta_acc, coef);
    SP_LAUNCH_OK("bn_bwd_finalize_kernel");
    return 0;
}

int sp_bn_act_bwd_apply(const float* gxh, int ldg, const float* x, int ldx, const float* coef, int N, int64_t vox, int C,
                        int G, int act, float alpha, float* out, int ldout, int accumulate, double* colsum, void* stream) {
    SP_REQUIRE(gxh && out, "sp_bn_act_bwd_apply: NULL pointer");
    SP_REQUIRE(x || (!coef && act == SP_ACT_NONE), "sp_bn_act_bwd_apply: x required");
    SP_REQUIRE(N > 0 && vox > 0 && C > 0 && ldg >= C && ldout >= C && G >= 1 && N % G == 0, "sp_bn_act_bwd_apply: bad shape");
    const int64_t rows = (int64_t)N * vox;
    const bool vec = (C % 4 == 0) && (ldg % 4 == 0) && (ldout % 4 == 0) && (!x || ldx % 4 == 0);
    if (colsum) SP_CUDA(cudaMemsetAsync(colsum, 0, sizeof(double) * C, sp_stream(stream)));
    if (vec) {
        const int C4 = C / 4;
        const int lanesQ = C4 < 256 ? C4 : 256;
        const int R = 256 / lanesQ;
        const int64_t rpg = (int64_t)(N / G) * vox;
        int64_t blocks = sp_cdiv(rpg, (int64_t)R * 4);       // ~4 rows per thread and group at least
        const int64_t cap = (int64_t)sp_num_sms() * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        bn_act_bwd_apply_vec_kernel<<<(int)blocks, 256, 0, sp_stream(stream)>>>(gxh, ldg, x, ldx, coef, rpg, C4, G, act, alpha, out,
                                                                             ldout, accumulate, colsum);
        SP_LAUNCH_OK("bn_act_bwd_apply_vec_kernel");
        return 0;
    }
    int64_t blocks = sp_cdiv(rows * C, 256 * 4);
    const int64_t cap = (int64_t)sp_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    bn_act_bwd_apply_kernel<<<(int)blocks, 256, 0, sp_stream(stream)>>>(gxh, ldg, x, ldx, coef, rows, (int64_t)(N / G) * vox, C, G,
                                                                     act, alpha, out, ldout, accumulate);
    SP_LAUNCH_OK("bn_act_bwd_apply_kernel");
    if (colsum) {   // scalar layout: column sums by the generic reduction over the tensor just written
        const int64_t rpb = sp_cdiv(rows, (int64_t)sp_num_sms() * 8);
        channel_sums_colonly_kernel<<<(int)sp_cdiv(rows, rpb), 256, 0, sp_stream(stream)>>>(out, ldout, rows, C, rpb, colsum);
        SP_LAUNCH_OK("channel_sums_colonly_kernel");
    }
    return 0;
}

int sp_bias_from_colsum(const double* colsum, int C, float* db, float beta, void* stream) {
    SP_REQUIRE(colsum && db && C > 0, "sp_bias_from_colsum: bad arguments");
    colsum_to_bias_kernel<<<(C + 255) / 256, 256, 0, sp_stream(stream)>>>(colsum, C, db, beta);
    SP_LAUNCH_OK("colsum_to_bias_kernel");
    return 0;
}

int sp_bn_grads_from_wgrad(const float* w, const float* dw_hat, const double* colsum, const double* tap_excl, const float* gamma,
                           const float* beta, int Co, int Ci, int k3, float* dw, float beta_dw, float* dgamma, float* dbeta,
                           float beta_acc, void* stream) {
    SP_REQUIRE(w && dw_hat && colsum && gamma && beta && dw, "sp_bn_grads_from_wgrad: NULL pointer");
    SP_REQUIRE(Co > 0 && Ci > 0 && k3 > 0, "sp_bn_grads_from_wgrad: bad extents %d %d %d", Co, Ci, k3);
    bn_grads_from_wgrad_kernel<<<Ci, 128, 0, sp_stream(stream)>>>(w, dw_hat, colsum, tap_excl, gamma, beta, Co, Ci, k3, dw, beta_dw,
                                                                   dgamma, dbeta, beta_acc);
    SP_LAUNCH_OK("bn_grads_from_wgrad_kernel");
    return 0;
}

int sp_border_tap_sums(const float* gz, int ldz, int N, int Do, int Ho, int Wo, int Co, int pd, int ph, int pw, double* excl,
                       void* stream) {
    SP_REQUIRE(gz && excl, "sp_border_tap_sums: NULL pointer");
    SP_REQUIRE(N > 0 && Do > 0 && Ho > 0 && Wo > 0 && Co > 0 && Co <= 256 && ldz >= Co, "sp_border_tap_sums: bad shape");
    SP_REQUIRE(pd >= 0 && pd <= 2 && ph >= 0 && ph <= 2 && pw >= 0 && pw <= 2, "sp_border_tap_sums: padding must be 0..2 (3x3x3, stride 1)");
    SP_CUDA(cudaMemsetAsync(excl, 0, sizeof(double) * 27 * Co * SP_TAP_EXCL_REPLICAS, sp_stream(stream)));
    if (pd == 0 && ph == 0 && pw == 0) return 0;
    int CoP2 = 1;
    while (CoP2 < Co) CoP2 *= 2;
    SP_REQUIRE((int64_t)N * Do * Ho < (1ll << 31), "sp_border_tap_sums: N * Do * Ho must fit 31 bits");
    const int64_t rows = (int64_t)N * Do * Ho;
    int64_t blocks = (int64_t)sp_num_sms() * 4;
    if (blocks > rows) blocks = rows;
    border_tap_sums_kernel<<<(int)blocks, 256, 0, sp_stream(stream)>>>(gz, ldz, N, Do, Ho, Wo, Co, CoP2, pd, ph, pw, Do + 2 - 2 * pd,
                                                                        Ho + 2 - 2 * ph, Wo + 2 - 2 * pw, excl);
    SP_LAUNCH_OK("border_tap_sums_kernel");
    fold_replicas_kernel<<<(27 * Co + 127) / 128, 128, 0, sp_stream(stream)>>>(excl, 27 * Co, SP_TAP_EXCL_REPLICAS);
    SP_LAUNCH_OK("fold_replicas_kernel");
    return 0;
}

}  // extern "C"
