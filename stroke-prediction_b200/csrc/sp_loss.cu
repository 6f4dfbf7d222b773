// sp_loss.cu — loss reductions and their gradients, plus the latent interpolation.
//   BatchDiceLoss.forward            common/metrics.py:16-28   (whole-batch soft Dice, eps = 1e-7)
//   hinge mean(abs(d) - d)           learner/CaeReconstructionLearner.py:59-62, CaeStepLearner.py:18-19
//   L1    mean(abs(a - b))           learner/CaeReconstructionLearner.py:68, CaePredictionLearner.py:53-55
//   Enc3D._interpolate               common/model/Cae3D.py:78-89
// HBM-bound streaming kernels: 128-bit loads, per-thread fp32 partials over short runs folded into fp64, warp
// shuffle + shared-memory block reduction, one fp64 atomicAdd per CTA and quantity.  Scalars stay on the device
// (no host synchronisation anywhere on the training step).
#include "sp_common.cuh"

namespace {

inline int red_grid(int64_t n) {
    int64_t b = sp_cdiv(n, 256 * 16);
    const int64_t cap = (int64_t)sp_num_sms() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}
inline int ew_grid(int64_t n) {
    int64_t b = sp_cdiv(n, 256 * 4);
    const int64_t cap = (int64_t)sp_num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

template <int NQ>
__device__ __forceinline__ void block_reduce_atomic(double (&v)[NQ], double* out) {
    __shared__ double sm[NQ][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        v[q] = sp_warp_sum(v[q]);
        if (lane == 0) sm[q][warp] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sm[threadIdx.x][w];
        atomicAdd(&out[threadIdx.x], s);
    }
}

// sums[0..2] = sum o*t, sum o*o, sum t*t
__global__ void __launch_bounds__(256)
dice_sums_kernel(const float* __restrict__ o, const float* __restrict__ t, int64_t n, double* __restrict__ sums) {
    double acc[3] = {0.0, 0.0, 0.0};
    const int64_t n4 = n / 4;
    const float4* o4 = reinterpret_cast<const float4*>(o);
    const float4* t4 = reinterpret_cast<const float4*>(t);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    while (i < n4) {
        float f0 = 0.f, f1 = 0.f, f2 = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i < n4) {
                const float4 a = sp_ldg_stream(o4 + i), b = sp_ldg_stream(t4 + i);
                f0 = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, f0))));
                f1 = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, f1))));
                f2 = fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, fmaf(b.w, b.w, f2))));
                i += stride;
            }
        }
        acc[0] += (double)f0; acc[1] += (double)f1; acc[2] += (double)f2;
    }
    // tail (n % 4 elements) handled by the first threads of CTA 0
    if (blockIdx.x == 0 && threadIdx.x < (n - n4 * 4)) {
        const float a = o[n4 * 4 + threadIdx.x], b = t[n4 * 4 + threadIdx.x];
        acc[0] += (double)a * b; acc[1] += (double)a * a; acc[2] += (double)b * b;
    }
    block_reduce_atomic<3>(acc, sums);
}

__global__ void dice_loss_kernel(const double* __restrict__ sums, float w, float eps, float* __restrict__ loss) {
    // the reference forms numerator / denominator in fp32 (metrics.py:25-27)
    const float inter = (float)sums[0];
    const float num = 2.f * inter + eps;
    const float den = (float)sums[1] + (float)sums[2] + eps;
    loss[0] = 1.f - w * (num / den);
}

__global__ void __launch_bounds__(256)
dice_bwd_kernel(const float* __restrict__ o, const float* __restrict__ t, int64_t n, const double* __restrict__ sums, float w,
                float eps, const float* __restrict__ gscale, float gmul, float* __restrict__ go, int accumulate) {
    const float num = 2.f * (float)sums[0] + eps;
    const float den = (float)sums[1] + (float)sums[2] + eps;
    const float g = (gscale ? gscale[0] : 1.f) * gmul;
    // d/do [1 - w*num/den] = -w * (2 t den - 2 o num) / den^2
    const float ka = -w * 2.f / den * g;          // multiplies t
    const float kb = w * 2.f * num / (den * den) * g;   // multiplies o
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float r = fmaf(ka, t[i], kb * o[i]);
        go[i] = accumulate ? go[i] + r : r;
    }
}

// mode 0: sum(|d| - d); mode 1: sum |d|, d = a - b
__global__ void __launch_bounds__(256)
absdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, int mode, double* __restrict__ sum) {
    double acc[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    while (i < n) {
        float f = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (i < n) {
                const float d = a[i] - b[i];
                f += (mode == 0) ? (fabsf(d) - d) : fabsf(d);
                i += stride;
            }
        }
        acc[0] += (double)f;
    }
    block_reduce_atomic<1>(acc, sum);
}
__global__ void mean_from_sum_kernel(const double* __restrict__ sum, double n, float* __restrict__ out) {
    out[0] = (float)(sum[0] / n);
}

__global__ void __launch_bounds__(256)
absdiff_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, int mode, const float* __restrict__ gscale,
                   float gmul, float* __restrict__ ga, int acc_a, float* __restrict__ gb, int acc_b) {
    const float g = (gscale ? gscale[0] : 1.f) * gmul / (float)n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);   // sign(0) = 0 like torch.abs backward
        const float r = g * ((mode == 0) ? (sgn - 1.f) : sgn);
        if (ga) ga[i] = acc_a ? ga[i] + r : r;
        if (gb) gb[i] = acc_b ? gb[i] - r : -r;
    }
}

// ---- latent interpolation ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
interp_fwd_kernel(const float* __restrict__ zc, const float* __restrict__ zp, const float* __restrict__ step, int64_t per,
                  int64_t total, float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float s = step[i / per];
        const float c = zc[i];
        out[i] = c + s * (zp[i] - c);   // association order of Cae3D.py:82-88 (no fma contraction of the subtraction)
    }
}

// one CTA row per sample for the dstep reduction; elementwise part in the same pass
__global__ void __launch_bounds__(256)
interp_bwd_kernel(const float* __restrict__ g, const float* __restrict__ zc, const float* __restrict__ zp,
                  const float* __restrict__ step, int64_t per, float* __restrict__ dzc, int acc_c, float* __restrict__ dzp,
                  int acc_p, double* __restrict__ dstep_acc) {
    const int b = blockIdx.y;
    const float s = step[b];
    const int64_t base = (int64_t)b * per;
    double acc[1] = {0.0};
    float f = 0.f;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < per; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + j;
        const float gv = g[i];
        if (dzc) dzc[i] = acc_c ? dzc[i] + gv * (1.f - s) : gv * (1.f - s);
        if (dzp) dzp[i] = acc_p ? dzp[i] + gv * s : gv * s;
        f = fmaf(gv, zp[i] - zc[i], f);
    }
    acc[0] = (double)f;
    if (dstep_acc) block_reduce_atomic<1>(acc, dstep_acc + b);
}
// ---- thresholded evaluation counts (metrics.py:31-47: result > 0.5 vs target > 0.5) ---------------------------------
// counts[0..3] = TP, FP, FN, TN as exact integers in doubles (a 28x128x128 batch is < 2^53 voxels by far)
__global__ void __launch_bounds__(256)
binary_counts_kernel(const float* __restrict__ r, const float* __restrict__ t, int64_t n, float thr, double* __restrict__ counts) {
    unsigned int c[4] = {0u, 0u, 0u, 0u};
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int iter = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool a = r[i] > thr, b = t[i] > thr;
        c[0] += (a && b) ? 1u : 0u;
        c[1] += (a && !b) ? 1u : 0u;
        c[2] += (!a && b) ? 1u : 0u;
        c[3] += (!a && !b) ? 1u : 0u;
        if (++iter == (1 << 20)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { acc[q] += (double)c[q]; c[q] = 0u; }
            iter = 0;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] += (double)c[q];
    block_reduce_atomic<4>(acc, counts);
}

__global__ void cast_d2f_kernel(const double* __restrict__ src, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}

}  // namespace

extern "C" {

int sp_dice_sums(const float* o, const float* t, int64_t n, double* sums, void* stream) {
    SP_REQUIRE(o && t && sums && n > 0, "sp_dice_sums: bad arguments");
    SP_REQUIRE(((uintptr_t)o % 16 == 0) && ((uintptr_t)t % 16 == 0), "sp_dice_sums: pointers must be 16-byte aligned");
    SP_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(double), sp_stream(stream)));
    dice_sums_kernel<<<red_grid(n / 4 + 1), 256, 0, sp_stream(stream)>>>(o, t, n, sums);
    SP_LAUNCH_OK("dice_sums_kernel");
    return 0;
}

int sp_dice_loss(const double* sums, float w, float eps, float* loss, void* stream) {
    SP_REQUIRE(sums && loss, "sp_dice_loss: NULL pointer");
    dice_loss_kernel<<<1, 1, 0, sp_stream(stream)>>>(sums, w, eps, loss);
    SP_LAUNCH_OK("dice_loss_kernel");
    return 0;
}

int sp_dice_bwd(const float* o, const float* t, int64_t n, const double* sums, float w, float eps, const float* gscale,
                float gmul, float* go, int accumulate, void* stream) {
    SP_REQUIRE(o && t && sums && go && n > 0, "sp_dice_bwd: bad arguments");
    dice_bwd_kernel<<<ew_grid(n), 256, 0, sp_stream(stream)>>>(o, t, n, sums, w, eps, gscale, gmul, go, accumulate);
    SP_LAUNCH_OK("dice_bwd_kernel");
    return 0;
}

int sp_absdiff_mean(const float* a, const float* b, int64_t n, int mode, double* sum_ws, float* out, void* stream) {
    SP_REQUIRE(a && b && sum_ws && out && n > 0, "sp_absdiff_mean: bad arguments");
    SP_REQUIRE(mode == 0 || mode == 1, "sp_absdiff_mean: mode must be 0 (hinge) or 1 (L1)");
    SP_CUDA(cudaMemsetAsync(sum_ws, 0, sizeof(double), sp_stream(stream)));
    absdiff_sum_kernel<<<red_grid(n), 256, 0, sp_stream(stream)>>>(a, b, n, mode, sum_ws);
    SP_LAUNCH_OK("absdiff_sum_kernel");
    mean_from_sum_kernel<<<1, 1, 0, sp_stream(stream)>>>(sum_ws, (double)n, out);
    SP_LAUNCH_OK("mean_from_sum_kernel");
    return 0;
}

int sp_binary_counts(const float* result, const float* target, int64_t n, float threshold, double* counts, void* stream) {
    SP_REQUIRE(result && target && counts && n > 0, "sp_binary_counts: bad arguments");
    SP_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(double), sp_stream(stream)));
    binary_counts_kernel<<<red_grid(n), 256, 0, sp_stream(stream)>>>(result, target, n, threshold, counts);
    SP_LAUNCH_OK("binary_counts_kernel");
    return 0;
}

int sp_absdiff_bwd(const float* a, const float* b, int64_t n, int mode, const float* gscale, float gmul, float* ga, int acc_a,
                   float* gb, int acc_b, void* stream) {
    SP_REQUIRE(a && b && n > 0 && (ga || gb), "sp_absdiff_bwd: bad arguments");
    SP_REQUIRE(mode == 0 || mode == 1, "sp_absdiff_bwd: mode must be 0 (hinge) or 1 (L1)");
    absdiff_bwd_kernel<<<ew_grid(n), 256, 0, sp_stream(stream)>>>(a, b, n, mode, gscale, gmul, ga, acc_a, gb, acc_b);
    SP_LAUNCH_OK("absdiff_bwd_kernel");
    return 0;
}

int sp_latent_interp_fwd(const float* zc, const float* zp, const float* step, int B, int64_t per_sample, float* out, void* stream) {
    SP_REQUIRE(zc && zp && step && out && B > 0 && per_sample > 0, "sp_latent_interp_fwd: bad arguments");
    const int64_t total = (int64_t)B * per_sample;
    interp_fwd_kernel<<<ew_grid(total), 256, 0, sp_stream(stream)>>>(zc, zp, step, per_sample, total, out);
    SP_LAUNCH_OK("interp_fwd_kernel");
    return 0;
}

int sp_latent_interp_bwd(const float* g, const float* zc, const float* zp, const float* step, int B, int64_t per_sample,
                         float* dzc, int acc_c, float* dzp, int acc_p, float* dstep, double* ws, void* stream) {
    SP_REQUIRE(g && zc && zp && step && B > 0 && per_sample > 0, "sp_latent_interp_bwd: bad arguments");
    SP_REQUIRE(!dstep || ws, "sp_latent_interp_bwd: dstep needs a workspace of B doubles");
    SP_REQUIRE(B <= 65535, "sp_latent_interp_bwd: batch too large");
    if (dstep) SP_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * B, sp_stream(stream)));
    int gx = (int)sp_cdiv(per_sample, 256 * 4);
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    dim3 grid(gx, B);
    interp_bwd_kernel<<<grid, 256, 0, sp_stream(stream)>>>(g, zc, zp, step, per_sample, dzc, acc_c, dzp, acc_p, dstep ? ws : nullptr);
    SP_LAUNCH_OK("interp_bwd_kernel");
    if (dstep) {
        cast_d2f_kernel<<<(B + 127) / 128, 128, 0, sp_stream(stream)>>>(ws, dstep, B);
        SP_LAUNCH_OK("cast_d2f_kernel");
    }
    return 0;
}

}  // extern "C"
