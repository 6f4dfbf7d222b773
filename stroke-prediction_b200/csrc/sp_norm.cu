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
    if (mode == 1)
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
                        int G, int act, float alpha, float* out, int ldout, int accumulate, void* stream) {
    SP_REQUIRE(gxh && out, "sp_bn_act_bwd_apply: NULL pointer");
    SP_REQUIRE(x || (!coef && act == SP_ACT_NONE), "sp_bn_act_bwd_apply: x required");
    SP_REQUIRE(N > 0 && vox > 0 && C > 0 && ldg >= C && ldout >= C && G >= 1 && N % G == 0, "sp_bn_act_bwd_apply: bad shape");
    const int64_t rows = (int64_t)N * vox;
    int64_t blocks = sp_cdiv(rows * C, 256 * 4);
    const int64_t cap = (int64_t)sp_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    bn_act_bwd_apply_kernel<<<(int)blocks, 256, 0, sp_stream(stream)>>>(gxh, ldg, x, ldx, coef, rows, (int64_t)(N / G) * vox, C, G,
                                                                     act, alpha, out, ldout, accumulate);
    SP_LAUNCH_OK("bn_act_bwd_apply_kernel");
    return 0;
}

}  // extern "C"
