// sp_optim.cu — fused multi-tensor Adam step.
// Replaces torch.optim.Adam.step() + optimizer.zero_grad() as called from learner/Learner.py:120-122 with the
// hyper-parameters of train_shape_reconstruction.py:11-13,40 / train_unet_segmentation.py:13-14,32 (L2-coupled
// weight decay, not AdamW) and the run-time beta1 schedule of CaeReconstructionLearner.py:28-40.
// Update rule = the installed-torch form pinned in SURVEY App. D:
//   g <- g*grad_scale + wd*p ; m <- m + (g - m)(1 - b1) ; v <- b2*v + (1 - b2) g^2
//   p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// One launch covers every parameter tensor: CTA b binary-searches the tensor whose [block_start, next) range holds
// b and processes SP_ADAM_CHUNK contiguous elements of it.  HBM-bound: 16 B read + 12 B written per parameter
// (+4 B when the gradient is cleared in the same pass).
#include "sp_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
adam_multi_kernel(const SpAdamTensor* __restrict__ table, int n_tensors, float step_size, float omb1, float beta2,
                  float omb2, float inv_sqrt_bc2, float eps, float wd, float grad_scale, int zero_grad) {
    // locate the tensor of this CTA
    int lo = 0, hi = n_tensors - 1;
    const int64_t b = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].block_start <= b) lo = mid; else hi = mid - 1;
    }
    const SpAdamTensor T = table[lo];
    const int64_t e0 = (b - T.block_start) * SP_ADAM_CHUNK;
    const int64_t e1 = (e0 + SP_ADAM_CHUNK < T.n) ? e0 + SP_ADAM_CHUNK : T.n;
    for (int64_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
        const float p = T.p[i];
        float g = T.g[i] * grad_scale;
        g = fmaf(wd, p, g);
        float m = T.m[i];
        float v = T.v[i];
        m = m + (g - m) * omb1;                             // torch: exp_avg.lerp_(grad, 1 - beta1)
        v = beta2 * v + omb2 * g * g;                       // torch: exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
        const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;  // torch: (sqrt(v) / sqrt(bc2)).add_(eps)
        T.p[i] = p - step_size * (m / denom);
        T.m[i] = m;
        T.v[i] = v;
        if (zero_grad) T.g[i] = 0.f;
    }
}

}  // namespace

extern "C" int sp_adam_multi(const SpAdamTensor* table, int n_tensors, int64_t total_blocks, double lr, double beta1,
                             double beta2, double eps, double weight_decay, int64_t step, double grad_scale,
                             int zero_grad, void* stream) {
    SP_REQUIRE(table && n_tensors > 0 && total_blocks > 0, "sp_adam_multi: empty table");
    SP_REQUIRE(step >= 1, "sp_adam_multi: step counts from 1");
    SP_REQUIRE(total_blocks < (1LL << 31), "sp_adam_multi: too many blocks");
    // hyper-parameters arrive as doubles: torch forms 1 - beta in double before rounding to fp32 (1 - 0.999f would be
    // off by 1.3e-5 relative)
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    const float step_size = (float)(lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    adam_multi_kernel<<<(unsigned)total_blocks, 256, 0, sp_stream(stream)>>>(
        table, n_tensors, step_size, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), inv_sqrt_bc2, (float)eps,
        (float)weight_decay, (float)grad_scale, zero_grad);
    SP_LAUNCH_OK("adam_multi_kernel");
    return 0;
}
