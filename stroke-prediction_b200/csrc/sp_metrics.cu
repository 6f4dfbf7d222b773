// sp_metrics.cu — surface-distance evaluation metrics (Hausdorff distance, average symmetric surface distance) and signed
// distance maps on the device: exact Euclidean distance transform by separable lower-envelope passes.
//
// Replaces the host round trip + MedPy of metrics.py:31-47 (`mpm.hd`, `mpm.assd`; MedPy==0.3.0, requirements.txt:2, not vendored:
// `__surface_distances` = border voxels by binary erosion with the connectivity-1 cross, scipy `distance_transform_edt` of the
// other mask's border complement, read at the own border voxels) and the `ndi.distance_transform_edt` calls of the SDM baseline
// (test_sdm_resampling.py:16-33).  HBM-bound integer work: squared distances are exact int32, one thread per lattice point,
// every pass reads its line through L1/L2 (the line of a thread block's neighbours is the same few cache lines).
#include "sp_common.cuh"
#include <math.h>

namespace {

constexpr int EDT_INF = 1 << 29;     // "no feature on this line yet"; INF + 2 * 65535^2 would overflow, extents are capped at 4096

// threshold + border extraction on a 4-axis lattice (n0, n1, n2, n3), dense, last axis contiguous.
// obj = (v > thr); border = obj and not eroded by the 2k-neighbour cross with an all-zero outside (scipy binary_erosion,
// border_value = 0).  all_border != 0: some axis of the caller's array has extent 1, so every set voxel has an outside
// neighbour along it and the erosion is empty (that is what MedPy computes on the reference's B x 1 x D x H x W arrays).
__global__ void border_kernel(const float* __restrict__ v, int n0, int n1, int n2, int n3, float thr, int all_border,
                              uint8_t* __restrict__ border, unsigned long long* __restrict__ count) {
    const int64_t total = (int64_t)n0 * n1 * n2 * n3;
    unsigned int local = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const bool obj = v[e] > thr;
        bool b = obj;
        if (obj && !all_border) {
            int64_t t = e;
            const int i3 = (int)(t % n3); t /= n3;
            const int i2 = (int)(t % n2); t /= n2;
            const int i1 = (int)(t % n1);
            const int i0 = (int)(t / n1);
            const int64_t s2 = n3, s1 = (int64_t)n2 * n3, s0 = (int64_t)n1 * n2 * n3;
            bool inner = true;      // every neighbour inside the array and set
            inner = inner && i3 > 0 && i3 < n3 - 1 && v[e - 1] > thr && v[e + 1] > thr;
            inner = inner && (n2 == 1 || (i2 > 0 && i2 < n2 - 1 && v[e - s2] > thr && v[e + s2] > thr));
            inner = inner && (n1 == 1 || (i1 > 0 && i1 < n1 - 1 && v[e - s1] > thr && v[e + s1] > thr));
            inner = inner && (n0 == 1 || (i0 > 0 && i0 < n0 - 1 && v[e - s0] > thr && v[e + s0] > thr));
            b = !inner;
        }
        border[e] = b ? 1 : 0;
        local += b ? 1u : 0u;
    }
    // block count -> one atomic (integer: order-independent, deterministic)
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ unsigned int ws[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) ws[warp] = local;
    __syncthreads();
    if (warp == 0) {
        unsigned int s = lane < (blockDim.x >> 5) ? ws[lane] : 0u;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s) atomicAdd(count, (unsigned long long)s);
    }
}

// feature mask from a float volume.  mode 0: !(v > thr)   1: (v > thr)   2: (v >= thr)
__global__ void feature_kernel(const float* __restrict__ v, int64_t total, float thr, int mode, uint8_t* __restrict__ feat) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const float x = v[e];
        feat[e] = (mode == 0 ? !(x > thr) : mode == 1 ? (x > thr) : (x >= thr)) ? 1 : 0;
    }
}

// One separable pass of the exact squared Euclidean distance transform along the axis of extent n and element stride s:
//   out[i] = min_j ( in[j] + (i - j)^2 )          (FIRST: in[j] = feat[j] ? 0 : INF)
// One thread per lattice point; for s > 1 the threads of a warp walk 32 neighbouring lines in lock step (coalesced), for
// s == 1 they share one line (broadcast).
template <bool FIRST>
__global__ void edt_pass_kernel(const void* __restrict__ in_, int64_t total, int n, int64_t s, int* __restrict__ out) {
    const uint8_t* feat = reinterpret_cast<const uint8_t*>(in_);
    const int* in = reinterpret_cast<const int*>(in_);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)((e / s) % n);
        const int64_t base = e - (int64_t)i * s;
        int best = EDT_INF;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const int dj = i - j;
            int v;
            if (FIRST) v = feat[base + (int64_t)j * s] ? 0 : EDT_INF;
            else v = in[base + (int64_t)j * s];
            v += dj * dj;
            best = v < best ? v : best;
        }
        out[e] = best < EDT_INF ? best : EDT_INF;
    }
}

// max / sum of sqrt(d2) over the voxels where sel != 0: per-block partials (fixed order -> deterministic), max by atomicMax
__global__ void surf_reduce_kernel(const int* __restrict__ d2, const uint8_t* __restrict__ sel, int64_t total,
                                   double* __restrict__ partial, int* __restrict__ maxd2) {
    double s = 0.0;
    int m = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        if (sel[e]) {
            const int d = d2[e];
            s += sqrt((double)d);
            m = d > m ? d : m;
        }
    }
    s = sp_warp_sum(s);
    for (int o = 16; o > 0; o >>= 1) {
        const int t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    __shared__ double ss[32];
    __shared__ int sm[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { ss[warp] = s; sm[warp] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        int mm = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += ss[w]; mm = sm[w] > mm ? sm[w] : mm; }
        partial[blockIdx.x] = t;
        if (mm > 0) atomicMax(maxd2, mm);
    }
}

// out[0] = hd, out[1] = assd, out[2] = asd(result -> target), out[3] = asd(target -> result), out[4] = hd(result -> target),
// out[5] = hd(target -> result), out[6] = result border voxels, out[7] = target border voxels.  Either mask empty: hd = assd = inf
// (metrics.py:36-37,43: the reference keeps numpy.Inf unless both masks have voxels).
__global__ void surf_final_kernel(const double* __restrict__ p1, const double* __restrict__ p2, int nblocks,
                                  const int* __restrict__ maxd2, const unsigned long long* __restrict__ counts,
                                  double* __restrict__ out) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < nblocks; ++i) { s1 += p1[i]; s2 += p2[i]; }
    const double n1 = (double)counts[0], n2 = (double)counts[1];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    out[6] = n1; out[7] = n2;
    if (n1 == 0.0 || n2 == 0.0) {
        out[0] = out[1] = out[2] = out[3] = out[4] = out[5] = inf;
        return;
    }
    const double h1 = sqrt((double)maxd2[0]), h2 = sqrt((double)maxd2[1]);
    const double a1 = s1 / n1, a2 = s2 / n2;
    out[0] = h1 > h2 ? h1 : h2;
    out[1] = 0.5 * (a1 + a2);          // numpy.mean((asd1, asd2))
    out[2] = a1; out[3] = a2; out[4] = h1; out[5] = h2;
}

// signed distance: out = sqrt(d_in) - sqrt(d_out)   (d_in: distance of object voxels to the background, d_out: of background
// voxels to the object); a mask with no object / no background has distance 0 on that side like scipy's EDT of an all-zero input...
__global__ void sdm_combine_kernel(const int* __restrict__ din, const int* __restrict__ dout, int64_t total, float sign,
                                   float* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const double a = din[e] >= EDT_INF ? 0.0 : sqrt((double)din[e]);
        const double b = dout[e] >= EDT_INF ? 0.0 : sqrt((double)dout[e]);
        out[e] = sign * (float)(a - b);
    }
}

int ew_grid(int64_t n, int threads = 256) {
    int64_t b = sp_cdiv(n, threads);
    const int64_t cap = (int64_t)sp_num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

constexpr int RED_BLOCKS = 592;     // 4 x 148

// exact squared EDT of `feat` (1 = feature) over the lattice (n0..n3): result in `a` (scratch `b`); both int32 [total]
int edt_run(const uint8_t* feat, int n0, int n1, int n2, int n3, int* a, int* b, cudaStream_t st) {
    const int64_t total = (int64_t)n0 * n1 * n2 * n3;
    const int g = ew_grid(total);
    int* cur = a;
    int* nxt = b;
    edt_pass_kernel<true><<<g, 256, 0, st>>>(feat, total, n3, 1, cur);
    SP_LAUNCH_OK("edt_pass_kernel<first>");
    const int ext[3] = {n2, n1, n0};
    const int64_t str[3] = {n3, (int64_t)n2 * n3, (int64_t)n1 * n2 * n3};
    for (int ax = 0; ax < 3; ++ax) {
        if (ext[ax] == 1) continue;
        edt_pass_kernel<false><<<g, 256, 0, st>>>(cur, total, ext[ax], str[ax], nxt);
        SP_LAUNCH_OK("edt_pass_kernel");
        int* t = cur; cur = nxt; nxt = t;
    }
    if (cur != a) SP_CUDA(cudaMemcpyAsync(a, cur, sizeof(int) * total, cudaMemcpyDeviceToDevice, st));
    return 0;
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" {

size_t sp_surface_distances_workspace_bytes(int64_t total) {
    if (total <= 0) return 0;
    // 2 border masks + 2 int32 lattices + 2 x RED_BLOCKS partial sums + counters
    return 2 * align256((size_t)total) + 2 * align256(sizeof(int) * (size_t)total) + align256(2 * RED_BLOCKS * sizeof(double)) + 256;
}

int sp_surface_distances(const float* result, const float* target, int n0, int n1, int n2, int n3, int all_border,
                         float threshold, double* out8, void* ws, size_t ws_bytes, void* stream) {
    SP_REQUIRE(result && target && out8 && ws, "sp_surface_distances: NULL pointer");
    SP_REQUIRE(n0 >= 1 && n1 >= 1 && n2 >= 1 && n3 >= 1 && n0 <= 4096 && n1 <= 4096 && n2 <= 4096 && n3 <= 4096,
               "sp_surface_distances: extents must be in [1, 4096], got %d %d %d %d", n0, n1, n2, n3);
    const int64_t total = (int64_t)n0 * n1 * n2 * n3;
    SP_REQUIRE(ws_bytes >= sp_surface_distances_workspace_bytes(total), "sp_surface_distances: workspace too small (%zu < %zu)",
               ws_bytes, sp_surface_distances_workspace_bytes(total));
    cudaStream_t st = sp_stream(stream);
    unsigned char* p = reinterpret_cast<unsigned char*>(ws);
    uint8_t* br = p;                      p += align256((size_t)total);
    uint8_t* bt = p;                      p += align256((size_t)total);
    int* da = reinterpret_cast<int*>(p);  p += align256(sizeof(int) * (size_t)total);
    int* db = reinterpret_cast<int*>(p);  p += align256(sizeof(int) * (size_t)total);
    double* part = reinterpret_cast<double*>(p); p += align256(2 * RED_BLOCKS * sizeof(double));
    unsigned long long* counts = reinterpret_cast<unsigned long long*>(p);      // [2]
    int* maxd2 = reinterpret_cast<int*>(counts + 2);                            // [2]
    SP_CUDA(cudaMemsetAsync(counts, 0, 256, st));
    const int g = ew_grid(total);
    border_kernel<<<g, 256, 0, st>>>(result, n0, n1, n2, n3, threshold, all_border, br, counts);
    SP_LAUNCH_OK("border_kernel");
    border_kernel<<<g, 256, 0, st>>>(target, n0, n1, n2, n3, threshold, all_border, bt, counts + 1);
    SP_LAUNCH_OK("border_kernel");
    // distances to the target's border, read at the result's border — and the other way round
    if (int e = edt_run(bt, n0, n1, n2, n3, da, db, st)) return e;
    surf_reduce_kernel<<<RED_BLOCKS, 256, 0, st>>>(da, br, total, part, maxd2);
    SP_LAUNCH_OK("surf_reduce_kernel");
    if (int e = edt_run(br, n0, n1, n2, n3, da, db, st)) return e;
    surf_reduce_kernel<<<RED_BLOCKS, 256, 0, st>>>(da, bt, total, part + RED_BLOCKS, maxd2 + 1);
    SP_LAUNCH_OK("surf_reduce_kernel");
    surf_final_kernel<<<1, 1, 0, st>>>(part, part + RED_BLOCKS, RED_BLOCKS, maxd2, counts, out8);
    SP_LAUNCH_OK("surf_final_kernel");
    return 0;
}

size_t sp_signed_distance_workspace_bytes(int64_t total) {
    if (total <= 0) return 0;
    return align256((size_t)total) + 3 * align256(sizeof(int) * (size_t)total) + 256;
}

int sp_signed_distance(const float* mask, int n0, int n1, int n2, int n3, float threshold, int outside_is_lt, float sign,
                       float* out, void* ws, size_t ws_bytes, void* stream) {
    SP_REQUIRE(mask && out && ws, "sp_signed_distance: NULL pointer");
    SP_REQUIRE(n0 >= 1 && n1 >= 1 && n2 >= 1 && n3 >= 1 && n0 <= 4096 && n1 <= 4096 && n2 <= 4096 && n3 <= 4096,
               "sp_signed_distance: extents must be in [1, 4096]");
    const int64_t total = (int64_t)n0 * n1 * n2 * n3;
    SP_REQUIRE(ws_bytes >= sp_signed_distance_workspace_bytes(total), "sp_signed_distance: workspace too small");
    cudaStream_t st = sp_stream(stream);
    unsigned char* p = reinterpret_cast<unsigned char*>(ws);
    uint8_t* feat = p;                     p += align256((size_t)total);
    int* din = reinterpret_cast<int*>(p);  p += align256(sizeof(int) * (size_t)total);
    int* dout = reinterpret_cast<int*>(p); p += align256(sizeof(int) * (size_t)total);
    int* tmp = reinterpret_cast<int*>(p);
    const int g = ew_grid(total);
    // distance of object voxels to the nearest background voxel: features = background
    feature_kernel<<<g, 256, 0, st>>>(mask, total, threshold, 0, feat);
    SP_LAUNCH_OK("feature_kernel");
    if (int e = edt_run(feat, n0, n1, n2, n3, din, tmp, st)) return e;
    // distance of background voxels to the nearest object voxel: features = object
    feature_kernel<<<g, 256, 0, st>>>(mask, total, threshold, outside_is_lt ? 2 : 1, feat);
    SP_LAUNCH_OK("feature_kernel");
    if (int e = edt_run(feat, n0, n1, n2, n3, dout, tmp, st)) return e;
    sdm_combine_kernel<<<g, 256, 0, st>>>(din, dout, total, sign, out);
    SP_LAUNCH_OK("sdm_combine_kernel");
    return 0;
}

}  // extern "C"
