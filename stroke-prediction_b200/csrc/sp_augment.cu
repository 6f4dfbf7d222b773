// sp_augment.cu — the augmentation / resampling transforms that feed the hot path, on the device.
//
// Replaces the per-sample scipy work of the reference's data pipeline (common/data.py): ElasticDeform (:313-351 —
// scipy.ndimage.gaussian_filter of three uniform noise fields, sigma 4 voxels, mode "constant"; map_coordinates order 1, mode
// "constant"), ResamplePlaneXY (:354-380 — per-slice scipy.ndimage.zoom, order 0 / 1), HemisphericFlip (:215-245), PadImages
// (:280-296).  All volumes here are dense [n_vol][D][H][W] (torch B x C x D x H x W after ToTensor, data.py:299-310: D = z,
// H = y, W = x = numpy axis 0).  HBM-bound streaming kernels, one thread per output element; fp64 where scipy computes in fp64
// (the noise fields and coordinates), fp32 images.
#include "sp_common.cuh"
#include <math.h>

namespace {

constexpr int MAX_RADIUS = 40;
struct GaussTaps {
    int radius;
    double w[2 * MAX_RADIUS + 1];
};

// one separable pass of scipy.ndimage.gaussian_filter (correlate1d, mode = "constant", cval = 0) along the axis of extent n
// and element stride s:  out[i] = sum_{t=-r..r} w[t + r] * in[i + t]   (0 outside the line)
__global__ void gauss_pass_kernel(const double* __restrict__ in, int64_t total, int n, int64_t s, GaussTaps taps,
                                  double* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)((e / s) % n);
        const int lo = i - taps.radius < 0 ? -i : -taps.radius;
        const int hi = i + taps.radius > n - 1 ? n - 1 - i : taps.radius;
        double acc = 0.0;
        for (int t = lo; t <= hi; ++t) acc += taps.w[t + taps.radius] * in[e + (int64_t)t * s];
        out[e] = acc;
    }
}

// ElasticDeform.elastic_transform (data.py:331-341).  With the numpy volume indexed [i = x][j = y][k = z] the reference samples
//   out[i,j,k] = image( i + dy[i,j,k], j + dx[i,j,k], k + dz[i,j,k] )      (np.meshgrid's default 'xy' indexing swaps the roles
// of the first two fields; it requires X == Y), dx / dy / dz = alpha * filtered field 1 / 2 / 3 (dz additionally * zscale).
// Here [k][j][i] = [d][h][w]: w-coordinate += alpha * f2, h-coordinate += alpha * f1, d-coordinate += alpha * zscale * f3.
// Interpolation: scipy map_coordinates order 1, mode "constant": any coordinate outside [0, n-1] gives cval = 0.
__global__ void elastic_warp_kernel(const float* __restrict__ img, const double* __restrict__ f1, const double* __restrict__ f2,
                                    const double* __restrict__ f3, int64_t total, int D, int H, int W, double alpha, double zscale,
                                    float* __restrict__ out) {
    const int64_t vox = (int64_t)D * H * W;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = e % vox;
        const float* base = img + (e - v);
        const int w = (int)(v % W), h = (int)((v / W) % H), d = (int)(v / ((int64_t)W * H));
        const double cw = (double)w + alpha * f2[e];
        const double ch = (double)h + alpha * f1[e];
        const double cd = (double)d + alpha * zscale * f3[e];
        float r = 0.f;
        if (cw >= 0.0 && cw <= (double)(W - 1) && ch >= 0.0 && ch <= (double)(H - 1) && cd >= 0.0 && cd <= (double)(D - 1)) {
            const int w0 = (int)floor(cw), h0 = (int)floor(ch), d0 = (int)floor(cd);
            const double tw = cw - w0, th = ch - h0, td = cd - d0;
            const int w1 = w0 + 1 < W ? w0 + 1 : w0, h1 = h0 + 1 < H ? h0 + 1 : h0, d1 = d0 + 1 < D ? d0 + 1 : d0;
            double acc = 0.0;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int dd = a ? d1 : d0;
                const double wd = a ? td : 1.0 - td;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int hh = b ? h1 : h0;
                    const double wh = b ? th : 1.0 - th;
                    const float* row = base + ((int64_t)dd * H + hh) * W;
                    acc += wd * wh * ((1.0 - tw) * (double)row[w0] + tw * (double)row[w1]);
                }
            }
            r = (float)acc;
        }
        out[e] = r;
    }
}

// ResamplePlaneXY (data.py:354-380): scipy.ndimage.zoom of every (y, x) slice, order 0 (nearest, floor(c + 0.5)) or 1 (linear),
// scipy's corner-aligned coordinate map c = o * (n_in - 1) / (n_out - 1).
__global__ void zoom_plane_kernel(const float* __restrict__ in, int64_t planes, int H, int W, int Ho, int Wo, int order,
                                  float* __restrict__ out) {
    const int64_t total = planes * Ho * Wo;
    const double zh = Ho > 1 ? (double)(H - 1) / (double)(Ho - 1) : 0.0;
    const double zw = Wo > 1 ? (double)(W - 1) / (double)(Wo - 1) : 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int ow = (int)(e % Wo), oh = (int)((e / Wo) % Ho);
        const int64_t p = e / ((int64_t)Wo * Ho);
        const float* src = in + p * H * W;
        const double ch = oh * zh, cw = ow * zw;
        float r;
        if (order == 0) {
            int ih = (int)floor(ch + 0.5), iw = (int)floor(cw + 0.5);
            ih = ih < 0 ? 0 : (ih > H - 1 ? H - 1 : ih);
            iw = iw < 0 ? 0 : (iw > W - 1 ? W - 1 : iw);
            r = src[(int64_t)ih * W + iw];
        } else {
            int h0 = (int)floor(ch), w0 = (int)floor(cw);
            h0 = h0 > H - 1 ? H - 1 : h0;
            w0 = w0 > W - 1 ? W - 1 : w0;
            const int h1 = h0 + 1 < H ? h0 + 1 : h0, w1 = w0 + 1 < W ? w0 + 1 : w0;
            const double th = ch - h0, tw = cw - w0;
            const double top = (1.0 - tw) * (double)src[(int64_t)h0 * W + w0] + tw * (double)src[(int64_t)h0 * W + w1];
            const double bot = (1.0 - tw) * (double)src[(int64_t)h1 * W + w0] + tw * (double)src[(int64_t)h1 * W + w1];
            r = (float)((1.0 - th) * top + th * bot);
        }
        out[e] = r;
    }
}

__global__ void flip_w_kernel(const float* __restrict__ in, int64_t rows, int W, float* __restrict__ out) {
    const int64_t total = rows * W;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(e % W);
        out[e] = in[e - w + (W - 1 - w)];
    }
}

__global__ void pad_kernel(const float* __restrict__ in, int64_t nvol, int D, int H, int W, int pd, int ph, int pw, float value,
                           float* __restrict__ out) {
    const int Do = D + 2 * pd, Ho = H + 2 * ph, Wo = W + 2 * pw;
    const int64_t total = nvol * Do * Ho * Wo;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(e % Wo) - pw, h = (int)((e / Wo) % Ho) - ph, d = (int)((e / ((int64_t)Wo * Ho)) % Do) - pd;
        const int64_t n = e / ((int64_t)Wo * Ho * Do);
        float r = value;
        if (w >= 0 && w < W && h >= 0 && h < H && d >= 0 && d < D) r = in[((n * D + d) * H + h) * (int64_t)W + w];
        out[e] = r;
    }
}

int aug_grid(int64_t n) {
    int64_t b = sp_cdiv(n, 256);
    const int64_t cap = (int64_t)sp_num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

extern "C" {

int sp_gauss3d(const double* in, int64_t nvol, int D, int H, int W, double sigma, double truncate, double* out, double* tmp,
               void* stream) {
    SP_REQUIRE(in && out && tmp && nvol > 0 && D > 0 && H > 0 && W > 0, "sp_gauss3d: bad arguments");
    SP_REQUIRE(sigma > 0.0 && truncate > 0.0, "sp_gauss3d: sigma and truncate must be positive");
    const int radius = (int)(truncate * sigma + 0.5);          // scipy: lw = int(truncate * sd + 0.5)
    SP_REQUIRE(radius <= MAX_RADIUS, "sp_gauss3d: kernel radius %d exceeds %d", radius, MAX_RADIUS);
    GaussTaps taps;
    taps.radius = radius;
    double sum = 0.0;
    for (int t = -radius; t <= radius; ++t) {                  // scipy _gaussian_kernel1d: exp(-0.5 / sigma^2 * x^2), normalised
        taps.w[t + radius] = exp(-0.5 / (sigma * sigma) * (double)t * (double)t);
        sum += taps.w[t + radius];
    }
    for (int t = 0; t <= 2 * radius; ++t) taps.w[t] /= sum;
    const int64_t total = nvol * D * H * W;
    cudaStream_t st = sp_stream(stream);
    const int g = aug_grid(total);
    // scipy filters axis 0, 1, 2 of the numpy (x, y, z) volume in turn = W, H, D here
    gauss_pass_kernel<<<g, 256, 0, st>>>(in, total, W, 1, taps, out);
    SP_LAUNCH_OK("gauss_pass_kernel");
    gauss_pass_kernel<<<g, 256, 0, st>>>(out, total, H, W, taps, tmp);
    SP_LAUNCH_OK("gauss_pass_kernel");
    gauss_pass_kernel<<<g, 256, 0, st>>>(tmp, total, D, (int64_t)H * W, taps, out);
    SP_LAUNCH_OK("gauss_pass_kernel");
    return 0;
}

int sp_elastic_warp(const float* img, const double* f1, const double* f2, const double* f3, int64_t nvol, int D, int H, int W,
                    double alpha, double zscale, float* out, void* stream) {
    SP_REQUIRE(img && f1 && f2 && f3 && out && nvol > 0 && D > 0 && H > 0 && W > 0, "sp_elastic_warp: bad arguments");
    SP_REQUIRE(img != out, "sp_elastic_warp: in-place operation is not possible");
    const int64_t total = nvol * D * H * W;
    elastic_warp_kernel<<<aug_grid(total), 256, 0, sp_stream(stream)>>>(img, f1, f2, f3, total, D, H, W, alpha, zscale, out);
    SP_LAUNCH_OK("elastic_warp_kernel");
    return 0;
}

int sp_zoom_plane_xy(const float* in, int64_t planes, int H, int W, int Ho, int Wo, int order, float* out, void* stream) {
    SP_REQUIRE(in && out && planes > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "sp_zoom_plane_xy: bad arguments");
    SP_REQUIRE(order == 0 || order == 1, "sp_zoom_plane_xy: order must be 0 (nearest) or 1 (linear)");
    zoom_plane_kernel<<<aug_grid(planes * Ho * Wo), 256, 0, sp_stream(stream)>>>(in, planes, H, W, Ho, Wo, order, out);
    SP_LAUNCH_OK("zoom_plane_kernel");
    return 0;
}

int sp_flip_w(const float* in, int64_t rows, int W, float* out, void* stream) {
    SP_REQUIRE(in && out && rows > 0 && W > 0 && in != out, "sp_flip_w: bad arguments");
    flip_w_kernel<<<aug_grid(rows * W), 256, 0, sp_stream(stream)>>>(in, rows, W, out);
    SP_LAUNCH_OK("flip_w_kernel");
    return 0;
}

int sp_pad_volume(const float* in, int64_t nvol, int D, int H, int W, int pd, int ph, int pw, float value, float* out, void* stream) {
    SP_REQUIRE(in && out && nvol > 0 && D > 0 && H > 0 && W > 0 && pd >= 0 && ph >= 0 && pw >= 0, "sp_pad_volume: bad arguments");
    const int64_t total = nvol * (D + 2 * pd) * (H + 2 * ph) * (int64_t)(W + 2 * pw);
    pad_kernel<<<aug_grid(total), 256, 0, sp_stream(stream)>>>(in, nvol, D, H, W, pd, ph, pw, value, out);
    SP_LAUNCH_OK("pad_kernel");
    return 0;
}

}  // extern "C"
