/*
 * stroke_b200.h — C-ABI of libstroke_b200.so (sm_100a kernels for the volumetric CAE / U-Net hot path).
 *
 * The reference (multimodallearning/stroke-prediction) has no FFI: every arithmetic op on its hot path is a call
 * into the torch 0.3.1 wheel from common/model/Cae3D.py, common/model/Unet3D.py, common/metrics.py and
 * learner/*.py.  Each entry point below therefore cites the reference call-site(s) (file:line, relative to the
 * reference root) whose torch op it replaces.
 *
 * Conventions
 *   - All tensors are fp32, dense NDHWC ("channels_last_3d"): element (n,d,h,w,c) lives at
 *     (((n*D + d)*H + h)*W + w)*ld + c, where ld (floats per voxel) >= C lets a kernel address a channel slice
 *     of a wider (concatenated) buffer.  Pointers are device pointers, 16-byte aligned.
 *   - The library never allocates tensor memory.  Workspaces are sized by the *_workspace_bytes query and
 *     passed in by the caller.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises the device,
 *     is re-entrant and CUDA-graph-capture safe.
 *   - Return value: 0 = success; negative = argument/shape error (nothing launched); positive = cudaError_t of
 *     the failed launch.  sp_last_error() returns a thread-local message.
 *   - "G" (statistic groups): a batch of N samples may be G independent BatchNorm calls stacked along N
 *     (the CAE runs its encoder 3x and decoder 4x per step with separate batch statistics,
 *     Cae3D.py:105-110,230-233).  Per-channel BN vectors are laid out [G][C]; sample n belongs to group n/(N/G).
 */
#ifndef STROKE_B200_H
#define STROKE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP_VERSION 100

/* activation codes (epilogues / derivative-from-output rules) */
enum SpAct {
    SP_ACT_NONE = 0,
    SP_ACT_ELU = 1,      /* nn.ELU(alpha)        Cae3D.py:42..216                      */
    SP_ACT_LEAKY = 2,    /* nn.LeakyReLU(alpha)  Unet3D.py:20,23,51                    */
    SP_ACT_SIGMOID = 3   /* nn.Sigmoid           Cae3D.py:136,219; Unet3D.py:53        */
};

/*
 * Geometry of one cubic convolution, stated for the *correlation* direction:
 *   O[n,o,co] = sum_{tap,ci} I[n, o*s - p + tap, ci] * W[co,ci,tap]
 * I ("I-side") is the larger tensor of a strided conv, O ("O-side") the smaller.
 *   nn.Conv3d forward            = sp_corr        (src = I-side, dst = O-side)
 *   nn.Conv3d dgrad              = sp_corrT       (src = O-side, dst = I-side)
 *   nn.ConvTranspose3d forward   = sp_corrT       (its weight (Cin_T,Cout_T,k,k,k) is exactly W[co][ci][tap])
 *   nn.ConvTranspose3d dgrad     = sp_corr
 *   weight gradient of either    = sp_wgrad
 */
typedef struct SpConvDesc {
    int32_t N;                      /* samples (all groups stacked)                                   */
    int32_t Di, Hi, Wi, Ci, ldi;    /* I-side extents, channels, floats per voxel                     */
    int32_t Do, Ho, Wo, Co, ldo;    /* O-side extents, channels, floats per voxel                     */
    int32_t k;                      /* cubic kernel extent: 1, 2 or 3                                 */
    int32_t s;                      /* stride (same on all axes): 1 or 2                              */
    int32_t pd, ph, pw;             /* zero padding of the I-side per axis (may exceed (k-1)/2)       */
    int32_t act;                    /* SpAct applied to dst after bias                                */
    float   alpha;                  /* ELU alpha / LeakyReLU slope                                    */
} SpConvDesc;

int         sp_version(void);
const char* sp_last_error(void);

/*
 * Tensor-core tiers of the 3x3x3 stride-1 layers (Cae3D.py:44,52,55,63,66,186-211; Unet3D.py:19,22): forward and dgrad with
 * 8..96 channels on either side (mode 4), weight gradients with 9..24 channels (wider I-sides as 16-channel slices).  Every
 * fp32 operand of the forward / dgrad kernels is staged as three bf16 terms (exact split), the products of order <= 2 are
 * accumulated by tcgen05.mma in TMEM (weight gradients: see sp_set_wgrad_tc_options):
 *   4 (default)  pipelined kernels, leading products in one accumulator per kd and corrections in separate columns: per-layer
 *                forward rel-L2 1.3e-7, weight gradient 5.7e-7 (an IEEE fp32 FFMA chain: 2.6e-7 / 1.7e-6 on the same data)
 *   0            tiers off (exact-fp32 FFMA tiers everywhere, weight gradients included)
 *   2 / 3        first-generation forward kernels, 8..16 channels only, one accumulator per output (rel-L2 ~5e-6 / ~1.4e-6),
 *                for A/B measurements
 * Packed weights depend on the mode: re-pack (sp_packed_weight_floats / sp_pack_weights) after changing it.
 */
int         sp_get_tc_terms(void);
int         sp_set_tc_terms(int terms);
/* Weight gradients on tcgen05 (Cae3D.py:44-66,186-211, 48, 59, 204; Unet3D.py:19,22).  Generation 2 (default): 3x3x3 stride-1
 * layers with 2..96 input and 9..64 output channels in ONE launch over all 16 x 16 channel slice pairs (sp_wgrad_tc4.cuh: one
 * M 128 x N 96 tcgen05.mma per (16 voxels, kd), two round-to-nearest bf16 terms per operand) and the stride-2 layers Conv3d k3 s2 p1 /
 * ConvTranspose3d k2 s2 with 9..32 channels (sp_wgrad_tc4s2.cuh).  Generation 1: the first-generation kernels (sp_wgrad_tc.cuh,
 * sp_wgrad_tc24.cuh: 27 M 64 x N 48 MMAs on three exact terms) for 9..24-channel stride-1 layers, FFMA tiers for everything else.
 * max_ctas > 0 caps the persistent grid (tests: several tile columns per CTA on small volumes), 0 = one CTA per SM. */
int         sp_set_wgrad_tc_options(int generation, int max_ctas);

/* ------------------------------------------------------------------------------------------------------------
 * Convolution family.  Replaces nn.Conv3d (Cae3D.py:41,44,48,52,55,59,63,66,70,74,126,128,132,186,189,197,200,
 * 208,211,215,218; Unet3D.py:19,22,50,52) and nn.ConvTranspose3d (Cae3D.py:178,182,193,204), forward + backward.
 *
 * Weights are consumed in a packed layout produced by sp_pack_weights from the torch layout W[co][ci][k^3]
 * (for ConvTranspose3d the torch layout (Cin_T,Cout_T,k,k,k) is already W[co][ci][tap] of the equivalent
 * correlation).  `which`: 0 = for sp_corr, 1 = for sp_corrT.
 *
 * Source prologue (optional, scale != NULL): v <- v*scale[g][c] + shift[g][c] applied to real voxels only, so
 * zero padding stays zero *after* BatchNorm (Cae3D.py:40-41 ordering BN -> padded conv).
 * Destination epilogue: + bias[c] (optional), then desc->act.
 * ---------------------------------------------------------------------------------------------------------- */
size_t sp_packed_weight_floats(const SpConvDesc* d, int which);
int    sp_pack_weights(const SpConvDesc* d, int which, const float* w_torch, float* w_packed, void* stream);

/* Scratch the correlation needs for this geometry (0 for the spatially tiled tiers; the wide, spatially tiny bottleneck
 * layers Cae3D.py:70,74,178,182 run as im2col / col2im + GEMM through a caller-provided workspace). */
size_t sp_conv_workspace_bytes(const SpConvDesc* d, int which);
int sp_corr (const SpConvDesc* d, const float* src_iside, const float* w_packed, const float* bias,
             const float* scale, const float* shift, int G, float* dst_oside, void* ws, size_t ws_bytes, void* stream);
int sp_corrT(const SpConvDesc* d, const float* src_oside, const float* w_packed, const float* bias,
             const float* scale, const float* shift, int G, float* dst_iside, void* ws, size_t ws_bytes, void* stream);

/* dW[co][ci][tap] (torch layout) = beta*dW + sum_{n,o} O'[n,o,co] * I'[n,o*s-p+tap,ci], where I' / O' are the
 * I-side / O-side tensors after their optional per-(group,channel) affine prologue (BatchNorm of the layer
 * input; zero padding stays zero).  Deterministic two-stage reduction through `ws`. */
size_t sp_wgrad_workspace_bytes(const SpConvDesc* d);
int    sp_wgrad(const SpConvDesc* d, const float* iside, const float* i_scale, const float* i_shift,
                const float* oside, const float* o_scale, const float* o_shift, int G,
                float* dw_torch, float beta, void* ws, size_t ws_bytes, void* stream);

/* db[c] = beta*db[c] + sum over rows of g[row*ld + c]   (bias gradient of Conv3d / ConvTranspose3d);
 * ws: C doubles of scratch */
int sp_bias_grad(const float* g, int64_t rows, int C, int ld, float* db, float beta, double* ws, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * BatchNorm3d (Cae3D.py:40..217, Unet3D.py:18,21), training and eval mode, forward statistics and backward.
 * The normalisation itself is never materialised: sp_bn_finalize emits scale/shift that the consuming
 * convolution applies while staging its input.
 * ---------------------------------------------------------------------------------------------------------- */
/* sums[g][c][0..1] (fp64) = sum x, sum x^2 over the group's voxels.  x: [N][vox][ld] */
int sp_bn_stats(const float* x, int N, int64_t vox, int C, int ld, int G, double* sums, void* stream);
/* training != 0: batch statistics (biased var for normalisation), running stats updated group after group with
 * momentum (unbiased var), num_batches_tracked += G.  training == 0: running stats.
 * Outputs scale/shift/mean/invstd as [G][C] floats. */
int sp_bn_finalize(const double* sums, int64_t count_per_group, int C, int G, const float* gamma,
                   const float* beta, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                   float momentum, float eps, int training, float* scale, float* shift, float* mean,
                   float* invstd, void* stream);
/* bsums[g][c][0..1] (fp64) = sum gxh, sum gxh*x  (gxh = gradient w.r.t. the BN output) */
int sp_bn_bwd_reduce(const float* gxh, int ldg, const float* x, int ldx, int N, int64_t vox, int C, int G,
                     double* bsums, void* stream);
/* dgamma/dbeta (accumulated with `beta_acc`, NULL to skip) and the coefficients of
 *   gx = A*((gxh - m1) - (x - mu)*k)        (coef layout [4][G][C]: A, m1, mu, k; same association as ATen so a
 *   large common mode of gxh cancels exactly) */
int sp_bn_bwd_finalize(const double* bsums, int64_t count_per_group, int C, int G, const float* gamma,
                       const float* mean, const float* invstd, int training, float* dgamma, float* dbeta,
                       float beta_acc, float* coef, void* stream);
/* out = A*((gxh - m1) - (x - mu)*k) * act'(x)   — BN backward apply fused with the backward of the activation that
 * produced x (derivative computed from the activation output).  coef == NULL: out = gxh*act'(x).
 * `accumulate` != 0: out += ...  */
int sp_bn_act_bwd_apply(const float* gxh, int ldg, const float* x, int ldx, const float* coef, int N,
                        int64_t vox, int C, int G, int act, float alpha, float* out, int ldout, int accumulate,
                        double* colsum, void* stream);
/* colsum (optional, C doubles): receives the per-channel column sums of the values written to `out`, i.e. the bias
 * gradient of the convolution whose output gradient `out` is; sp_bias_from_colsum turns it into db = beta*db + colsum */
int sp_bias_from_colsum(const double* colsum, int C, float* db, float beta, void* stream);
/* BatchNorm parameter gradients of a network's FIRST unit (BatchNorm3d on the data -> Conv3d without padding: Unet3D.py:16-18
 * block1, whose input needs no gradient, Learner.py:121) from the weight gradient instead of a dgrad + reduction:
 * dw_hat = sp_wgrad taken with scale = invstd, shift = -mean*invstd (the normalised input), colsum = the fp64 bias gradient;
 *   dgamma[ci] = beta_acc*dgamma + sum_{co,tap} w*dw_hat,  dbeta[ci] = beta_acc*dbeta + sum_{co,tap} w*colsum[co],
 *   dw = beta_dw*dw + gamma[ci]*dw_hat + beta[ci]*colsum[co].   w, dw_hat, dw: torch layout [Co][Ci][k3].
 * tap_excl (optional, [Co][k3] doubles): with zero padding (applied after BatchNorm, Cae3D.py:40-41) colsum[co] is replaced by
 * colsum[co] - tap_excl[co][tap], the sum over the output voxels whose tap reads a real input voxel. */
int sp_bn_grads_from_wgrad(const float* w, const float* dw_hat, const double* colsum, const double* tap_excl, const float* gamma,
                           const float* beta, int Co, int Ci, int k3, float* dw, float beta_dw, float* dgamma, float* dbeta,
                           float beta_acc, void* stream);
/* excl[co][tap] = sum of gz[co, v] over the output voxels v (N x Do x Ho x Wo, channel stride ldz) of a 3x3x3 stride-1 convolution
 * with padding (pd, ph, pw) in 0..2 whose tap (kd, kh, kw) reads zero padding; only border voxels are read.  `excl` must hold
 * SP_TAP_EXCL_REPLICAS * Co * 27 doubles (scratch copies that spread the atomics); the result is its first Co * 27 entries. */
#define SP_TAP_EXCL_REPLICAS 16
int sp_border_tap_sums(const float* gz, int ldz, int N, int Do, int Ho, int Wo, int Co, int pd, int ph, int pw, double* excl,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Resampling ops of the U-Net (Unet3D.py:39,41 MaxPool3d(2,2); :44,46 Upsample(x2, trilinear); :6-11,66-67,71-72
 * crop + cat) and layout glue.
 * ---------------------------------------------------------------------------------------------------------- */
int sp_maxpool2_fwd(const float* x, int N, int D, int H, int W, int C, int ldx, float* y, int ldy, void* stream);
/* gx (dense, ld = C) = route gy to the first maximum of each 2x2x2 window in d->h->w scan order, 0 elsewhere */
int sp_maxpool2_bwd(const float* x, const float* y, const float* gy, int N, int D, int H, int W, int C,
                    float* gx, void* stream);
int sp_upsample2_fwd(const float* x, int N, int D, int H, int W, int C, int ldx, float* y, int ldy,
                     int align_corners, void* stream);
int sp_upsample2_bwd(const float* gy, int ldgy, int N, int D, int H, int W, int C, float* gx, int ldgx,
                     int align_corners, void* stream);
/* dst[n, d, h, w, 0:C] (ld = ldd) (+)= src[n, d+od, h+oh, w+ow, 0:C] (ld = lds) over the dst extents;
 * used for centre-crop into a concat slice (forward) */
int sp_crop_copy(const float* src, int Ds, int Hs, int Ws, int lds, float* dst, int Dd, int Hd, int Wd, int ldd,
                 int N, int C, int od, int oh, int ow, void* stream);
/* big[n, d+od, h+oh, w+ow, 0:C] += small[n,d,h,w,0:C]  (gradient of the centre-crop, accumulated in place) */
int sp_crop_add(float* big, int Db, int Hb, int Wb, int ldb, const float* small, int Ds, int Hs, int Ws, int lds,
                int N, int C, int od, int oh, int ow, void* stream);
/* dense NCDHW <-> NDHWC */
int sp_ncdhw_to_ndhwc(const float* src, float* dst, int N, int C, int64_t vox, void* stream);
int sp_ndhwc_to_ncdhw(const float* src, float* dst, int N, int C, int64_t vox, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Losses.  BatchDiceLoss (metrics.py:16-28), hinge mean(abs(d)-d) and L1 mean(abs(a-b))
 * (CaeReconstructionLearner.py:59-62,68; CaeStepLearner.py:18-19; CaePredictionLearner.py:46-55).
 * Reductions accumulate in fp64 on the device; no host synchronisation.
 * ---------------------------------------------------------------------------------------------------------- */
/* sums[0..2] (fp64, must be zeroed by the call itself) = sum o*t, sum o*o, sum t*t over n elements */
int sp_dice_sums(const float* o, const float* t, int64_t n, double* sums, void* stream);
/* loss[0] = 1 - w*(2*sums[0]+eps)/(sums[1]+sums[2]+eps) */
int sp_dice_loss(const double* sums, float w, float eps, float* loss, void* stream);
/* go[i] (+)= gscale[0]*gmul * ( -w*(2*t*den - 2*o*num)/den^2 ) */
int sp_dice_bwd(const float* o, const float* t, int64_t n, const double* sums, float w, float eps,
                const float* gscale, float gmul, float* go, int accumulate, void* stream);
/* mode 0: hinge  out = mean(|a-b| - (a-b));  mode 1: L1  out = mean(|a-b|) */
int sp_absdiff_mean(const float* a, const float* b, int64_t n, int mode, double* sum_ws, float* out, void* stream);
/* ga (+)= g*gmul/n * dmode(a-b), gb (+)= -(same); either may be NULL.  sign(0) = 0. */
int sp_absdiff_bwd(const float* a, const float* b, int64_t n, int mode, const float* gscale, float gmul,
                   float* ga, int acc_a, float* gb, int acc_b, void* stream);

/* Evaluation counts of the thresholded masks (replaces the host round trip + medpy dc / precision / sensitivity /
 * specificity of metrics.py:31-47,49-62): counts[0..3] = TP, FP, FN, TN of (result > threshold) against
 * (target > threshold) over n contiguous floats. */
int sp_binary_counts(const float* result, const float* target, int64_t n, float threshold, double* counts,
                     void* stream);

/* Surface distances of the thresholded masks (replaces `mpm.hd` / `mpm.assd`, metrics.py:43-45; MedPy==0.3.0
 * `__surface_distances`: border = mask XOR binary_erosion(mask, connectivity-1 cross, outside = 0); distances = exact
 * Euclidean distance transform of the OTHER mask's border complement, read at the own border voxels, unit voxel spacing).
 * The arrays are a dense (n0, n1, n2, n3) lattice, n3 contiguous; an extent-1 axis may be dropped by the caller, who then
 * passes all_border != 0: with an extent-1 axis every set voxel has an outside neighbour, the erosion is empty and the whole
 * object is "border" — exactly what MedPy computes on the reference's B x 1 x D x H x W batches (the batch axis is a lattice
 * axis with unit spacing there, and is one here).
 * out8: hd, assd, asd(result->target), asd(target->result), hd(result->target), hd(target->result), border voxel counts of
 * result and target; hd = assd = +inf when either mask is empty (metrics.py:36-37,43). */
size_t sp_surface_distances_workspace_bytes(int64_t total);
int sp_surface_distances(const float* result, const float* target, int n0, int n1, int n2, int n3, int all_border,
                         float threshold, double* out8, void* ws, size_t ws_bytes, void* stream);

/* Signed distance map of a thresholded mask (the SDM baseline, test_sdm_resampling.py:16-33):
 *   out = sign * ( edt(v > thr) - edt(outside) ),  outside = (v < thr) when outside_is_lt != 0 (the penumbra form, :17-18)
 *   or !(v > thr) (the core form, :31-32, with sign = -1); edt(m) = distance of every voxel of m to the nearest voxel outside m
 *   (scipy.ndimage.distance_transform_edt), exact, unit spacing, over the (n0, n1, n2, n3) lattice. */
size_t sp_signed_distance_workspace_bytes(int64_t total);
int sp_signed_distance(const float* mask, int n0, int n1, int n2, int n3, float threshold, int outside_is_lt, float sign,
                       float* out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Augmentation / resampling transforms in front of the hot path (common/data.py:215-380), on dense [n_vol][D][H][W]
 * volumes (torch B x C x D x H x W after ToTensor, data.py:299-310: D = z, H = y, W = x).
 * ---------------------------------------------------------------------------------------------------------- */
/* scipy.ndimage.gaussian_filter(field, sigma, mode="constant", cval=0, truncate) of n_vol fp64 volumes (ElasticDeform's
 * noise smoothing, data.py:336-338); out and tmp: n_vol*D*H*W doubles each */
int sp_gauss3d(const double* in, int64_t nvol, int D, int H, int W, double sigma, double truncate, double* out, double* tmp,
               void* stream);
/* ElasticDeform.elastic_transform (data.py:331-341): out[d,h,w] = img( d + alpha*zscale*f3, h + alpha*f1, w + alpha*f2 ) by
 * linear interpolation, 0 outside the volume (map_coordinates order 1, mode "constant"); f1..f3 = the smoothed noise fields in
 * the reference's draw order */
int sp_elastic_warp(const float* img, const double* f1, const double* f2, const double* f3, int64_t nvol, int D, int H, int W,
                    double alpha, double zscale, float* out, void* stream);
/* ResamplePlaneXY (data.py:354-380): scipy.ndimage.zoom of every (H, W) plane to (Ho, Wo), order 0 (nearest) or 1 (linear) */
int sp_zoom_plane_xy(const float* in, int64_t planes, int H, int W, int Ho, int Wo, int order, float* out, void* stream);
/* HemisphericFlip (data.py:215-245): mirror along x (= W) */
int sp_flip_w(const float* in, int64_t rows, int W, float* out, void* stream);
/* PadImages (data.py:280-296): constant border of (pd, ph, pw) voxels */
int sp_pad_volume(const float* in, int64_t nvol, int D, int H, int W, int pd, int ph, int pw, float value, float* out,
                  void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Latent interpolation z_c + s*(z_p - z_c), s per sample (Cae3D.py:78-89).
 * ---------------------------------------------------------------------------------------------------------- */
int sp_latent_interp_fwd(const float* zc, const float* zp, const float* step, int B, int64_t per_sample,
                         float* out, void* stream);
/* dzc (+)= g*(1-s), dzp (+)= g*s, dstep[b] = sum g*(zp-zc)  (any output may be NULL; ws: B doubles when
 * dstep is requested) */
int sp_latent_interp_bwd(const float* g, const float* zc, const float* zp, const float* step, int B,
                         int64_t per_sample, float* dzc, int acc_c, float* dzp, int acc_p, float* dstep,
                         double* ws, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Fused multi-tensor Adam step (torch.optim.Adam as configured at train_shape_reconstruction.py:40,
 * train_unet_segmentation.py:32; beta1 schedule CaeReconstructionLearner.py:28-40):
 *   g += wd*p; m += (g-m)(1-b1); v = b2*v + (1-b2) g^2; p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * `table` is a device array of SpAdamTensor; one launch updates all tensors.  grad_scale multiplies g first
 * (1/world_size after a sum all-reduce).  Hyper-parameters are doubles because torch forms 1 - beta in double.  zero_grad != 0 clears g afterwards (optimizer.zero_grad, Learner.py:120).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct SpAdamTensor {
    float*  p;
    float*  g;
    float*  m;
    float*  v;
    int64_t n;
    int64_t block_start;   /* first CTA index that works on this tensor (prefix sum of ceil(n/SP_ADAM_CHUNK)) */
} SpAdamTensor;
#define SP_ADAM_CHUNK 4096
int sp_adam_multi(const SpAdamTensor* table, int n_tensors, int64_t total_blocks, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int64_t step, double grad_scale, int zero_grad,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STROKE_B200_H */
