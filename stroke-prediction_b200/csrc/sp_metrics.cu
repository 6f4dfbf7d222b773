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
                              uint8_t* __restrict__ border, unsigned long long* __restrict__ count, int* __restrict__ bb) {
    const int64_t total = (int64_t)n0 * n1 * n2 * n3;
    unsigned int local = 0;
    int lo[4] = {1 << 30, 1 << 30, 1 << 30, 1 << 30}, hi[4] = {-1, -1, -1, -1};
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const bool obj = v[e] > thr;
        bool b = obj;
        if (obj) {
            int64_t t = e;
            const int i3 = (int)(t % n3); t /= n3;
            const int i2 = (int)(t % n2); t /= n2;
            const int i1 = (int)(t % n1);
            const int i0 = (int)(t / n1);
            // bounding box of the union of both masks (this kernel runs once per mask on the same box)
            lo[0] = i0 < lo[0] ? i0 : lo[0]; hi[0] = i0 > hi[0] ? i0 : hi[0];
            lo[1] = i1 < lo[1] ? i1 : lo[1]; hi[1] = i1 > hi[1] ? i1 : hi[1];
            lo[2] = i2 < lo[2] ? i2 : lo[2]; hi[2] = i2 > hi[2] ? i2 : hi[2];
            lo[3] = i3 < lo[3] ? i3 : lo[3]; hi[3] = i3 > hi[3] ? i3 : hi[3];
          if (!all_border) {
            const int64_t s2 = n3, s1 = (int64_t)n2 * n3, s0 = (int64_t)n1 * n2 * n3;
            bool inner = true;      // every neighbour inside the array and set
            inner = inner && i3 > 0 && i3 < n3 - 1 && v[e - 1] > thr && v[e + 1] > thr;
            inner = inner && (n2 == 1 || (i2 > 0 && i2 < n2 - 1 && v[e - s2] > thr && v[e + s2] > thr));
            inner = inner && (n1 == 1 || (i1 > 0 && i1 < n1 - 1 && v[e - s1] > thr && v[e + s1] > thr));
            inner = inner && (n0 == 1 || (i0 > 0 && i0 < n0 - 1 && v[e - s0] > thr && v[e + s0] > thr));
            b = !inner;
          }
        }
        border[e] = b ? 1 : 0;
        local += b ? 1u : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            const int l = __shfl_xor_sync(0xffffffffu, lo[k], o), h = __shfl_xor_sync(0xffffffffu, hi[k], o);
            lo[k] = l < lo[k] ? l : lo[k];
            hi[k] = h > hi[k] ? h : hi[k];
        }
        if ((threadIdx.x & 31) == 0 && hi[k] >= 0) { atomicMin(&bb[2 * k], lo[k]); atomicMax(&bb[2 * k + 1], hi[k]); }
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

// Bounding box (per lattice axis) of the union of both thresholded masks, gathered by border_kernel: bb[2a] = lowest,
// bb[2a+1] = highest index on axis a (initialised to extent / -1).  Every distance the surface metrics read is taken AT a voxel of one mask TO a voxel of the
// other, so the transform only has to be exact inside this box: the passes skip lattice points outside it and scan their lines
// only across it.  The box lives in device memory — no host synchronisation — and typically holds 20 - 30 % of the lattice.
__global__ void bbox_init_kernel(int* bb, int n0, int n1, int n2, int n3) {
    if (threadIdx.x < 4) {
        const int n[4] = {n0, n1, n2, n3};
        bb[2 * threadIdx.x] = n[threadIdx.x];
        bb[2 * threadIdx.x + 1] = -1;
    }
}

// One separable pass of the exact squared Euclidean distance transform along the axis of extent n and element stride s:
//   out[i] = min_j ( in[j] + (i - j)^2 )          (FIRST: in[j] = feat[j] ? 0 : INF)
// One thread per lattice point; for s > 1 the threads of a warp walk 32 neighbouring lines in lock step (coalesced), for
// s == 1 they share one line (broadcast).
// First pass (contiguous axis): squared distance to the nearest feature of the same line.  One warp per line: the line's feature
// bits are gathered with ballots (32 positions each), every lane then finds the nearest set bit to the left and to the right of
// its positions with clz / ffs on the chunk masks — O(1) work per lattice point instead of a scan over the line.
__global__ void edt_first_kernel(const uint8_t* __restrict__ feat, int n0, int n1, int n2, int n3, const int* __restrict__ bb,
                                 int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t lines = (int64_t)n0 * n1 * n2;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int lo3 = bb ? bb[6] : 0, hi3 = bb ? bb[7] : n3 - 1;
    if (hi3 < lo3) return;
    const int c_lo = lo3 >> 5, c_hi = hi3 >> 5;                       // 32-wide chunks that intersect the box
    for (int64_t line = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; line < lines; line += warps) {
        if (bb) {
            int64_t t = line;
            const int c2 = (int)(t % n2); t /= n2;
            const int c1 = (int)(t % n1);
            const int c0 = (int)(t / n1);
            if (c0 < bb[0] || c0 > bb[1] || c1 < bb[2] || c1 > bb[3] || c2 < bb[4] || c2 > bb[5]) continue;
        }
        const uint8_t* f = feat + line * n3;
        int* o = out + line * n3;
        uint32_t mask[8];                                             // extents are capped at 256 for this kernel
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int i = c * 32 + lane;
            mask[c] = (c >= c_lo && c <= c_hi) ? __ballot_sync(0xffffffffu, i >= lo3 && i <= hi3 && f[i] != 0) : 0u;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int i = c * 32 + lane;
            if (c < c_lo || c > c_hi || i < lo3 || i > hi3) continue;
            int best = EDT_INF;
            // nearest feature at or left of i
            uint32_t m = mask[c] & (0xffffffffu >> (31 - lane));
            int pos = m ? c * 32 + 31 - __clz(m) : -1;
#pragma unroll
            for (int cc = 7; cc >= 0; --cc)
                if (cc < c && pos < 0 && mask[cc]) pos = cc * 32 + 31 - __clz(mask[cc]);
            if (pos >= 0) best = (i - pos) * (i - pos);
            // nearest feature right of i
            m = lane < 31 ? mask[c] & (0xffffffffu << (lane + 1)) : 0u;
            pos = m ? c * 32 + __ffs(m) - 1 : -1;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
                if (cc > c && pos < 0 && mask[cc]) pos = cc * 32 + __ffs(mask[cc]) - 1;
            if (pos >= 0) { const int dd = (pos - i) * (pos - i); best = dd < best ? dd : best; }
            o[i] = best;
        }
    }
}

// Later passes (axis stride >= n3), shared-memory form for extents <= 256: a CTA takes a tile of 32 neighbouring lines (32
// consecutive positions of the contiguous axis), loads the tile once (coalesced rows of 128 bytes) and evaluates the lower envelope
// from shared memory (column = lane: conflict free).  The global-memory form below re-reads its line from L2 for every candidate.
__global__ void edt_tile_kernel(const int* __restrict__ in, int n0, int n1, int n2, int n3, int axis, const int* __restrict__ bb,
                                int* __restrict__ out) {
    extern __shared__ int tile[];                        // [ext][32]
    const int ext[4] = {n0, n1, n2, n3};
    const int64_t str[4] = {(int64_t)n1 * n2 * n3, (int64_t)n2 * n3, (int64_t)n3, 1};
    int lo[4], hi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { lo[k] = bb ? bb[2 * k] : 0; hi[k] = bb ? bb[2 * k + 1] : ext[k] - 1; }
    if (hi[0] < lo[0]) return;
    // tiles: the two lattice axes other than `axis` and 3 (call them p, q) x chunks of 32 along axis 3, all restricted to the box
    int pa = -1, qa = -1;
    for (int k = 0; k < 3; ++k)
        if (k != axis) { if (pa < 0) pa = k; else qa = k; }
    const int np = hi[pa] - lo[pa] + 1, nq = hi[qa] - lo[qa] + 1;
    const int w0 = lo[3] & ~31, nwc = (hi[3] - w0) / 32 + 1;
    const int64_t ntiles = (int64_t)np * nq * nwc;
    const int na = hi[axis] - lo[axis] + 1;
    const int64_t sa = str[axis];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nty = blockDim.x >> 5;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int wc = (int)(t % nwc);
        const int iq = (int)((t / nwc) % nq), ip = (int)(t / ((int64_t)nwc * nq));
        const int w = w0 + wc * 32 + tx;
        const bool okw = w >= lo[3] && w <= hi[3];
        const int64_t base = (int64_t)(lo[pa] + ip) * str[pa] + (int64_t)(lo[qa] + iq) * str[qa] + w;
        __syncthreads();                                 // the previous tile has been consumed
        for (int i = ty; i < na; i += nty) tile[i * 32 + tx] = okw ? in[base + (int64_t)(lo[axis] + i) * sa] : EDT_INF;
        __syncthreads();
        if (!okw) continue;
        for (int i = ty; i < na; i += nty) {
            int best = tile[i * 32 + tx];
            const int down = i, up = na - 1 - i;
            const int far = down > up ? down : up;
            for (int dj = 1; dj <= far && dj * dj < best; ++dj) {
                const int q = dj * dj;
                if (dj <= down) { const int v = tile[(i - dj) * 32 + tx] + q; best = v < best ? v : best; }
                if (dj <= up) { const int v = tile[(i + dj) * 32 + tx] + q; best = v < best ? v : best; }
            }
            out[base + (int64_t)(lo[axis] + i) * sa] = best < EDT_INF ? best : EDT_INF;
        }
    }
}

template <bool FIRST>
__global__ void edt_pass_kernel(const void* __restrict__ in_, int n0, int n1, int n2, int n3, int axis, const int* __restrict__ bb,
                                int* __restrict__ out) {
    const uint8_t* feat = reinterpret_cast<const uint8_t*>(in_);
    const int* in = reinterpret_cast<const int*>(in_);
    const int64_t total = (int64_t)n0 * n1 * n2 * n3;
    const int ext[4] = {n0, n1, n2, n3};
    const int64_t str[4] = {(int64_t)n1 * n2 * n3, (int64_t)n2 * n3, (int64_t)n3, 1};
    int lo[4], hi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { lo[k] = bb ? bb[2 * k] : 0; hi[k] = bb ? bb[2 * k + 1] : ext[k] - 1; }
    const int64_t s = str[axis];
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = e;
        int c[4];
        c[3] = (int)(t % n3); t /= n3;
        c[2] = (int)(t % n2); t /= n2;
        c[1] = (int)(t % n1);
        c[0] = (int)(t / n1);
        if (c[0] < lo[0] || c[0] > hi[0] || c[1] < lo[1] || c[1] > hi[1] || c[2] < lo[2] || c[2] > hi[2] || c[3] < lo[3] || c[3] > hi[3])
            continue;                                   // outside the box: never read by a later pass or by the reduction
        const int i = c[axis];
        if (FIRST) {
            const int64_t base = e - (int64_t)i * s;
            int best = EDT_INF;
#pragma unroll 4
            for (int j = lo[axis]; j <= hi[axis]; ++j) {
                const int dj = i - j;
                const int v = (feat[base + (int64_t)j * s] ? 0 : EDT_INF) + dj * dj;
                best = v < best ? v : best;
            }
            out[e] = best < EDT_INF ? best : EDT_INF;
        } else {
            // lower envelope min_j in[j] + (i - j)^2, searched outwards from j = i: a candidate at distance dj can only win
            // while dj^2 < best, so lattice points near a feature stop after a few steps
            const int* p = in + e;
            int best = *p;
            const int down = i - lo[axis], up = hi[axis] - i;
            const int far = down > up ? down : up;
            for (int dj = 1; dj <= far && dj * dj < best; ++dj) {
                const int q = dj * dj;
                if (dj <= down) { const int v = p[-(int64_t)dj * s] + q; best = v < best ? v : best; }
                if (dj <= up) { const int v = p[(int64_t)dj * s] + q; best = v < best ? v : best; }
            }
            out[e] = best < EDT_INF ? best : EDT_INF;
        }
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
    // one warp, fixed association: lane l sums partials l, l + 32, ... in order, then a shuffle tree
    const int lane = threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < nblocks; i += 32) { s1 += p1[i]; s2 += p2[i]; }
    s1 = sp_warp_sum(s1);
    s2 = sp_warp_sum(s2);
    if (lane != 0) return;
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

// exact squared EDT of `feat` (1 = feature) over the lattice (n0..n3), restricted to the box `bb` (NULL: whole lattice): result
// in `a` (scratch `b`); both int32 [total]
int edt_run(const uint8_t* feat, int n0, int n1, int n2, int n3, const int* bb, int* a, int* b, cudaStream_t st) {
    const int64_t total = (int64_t)n0 * n1 * n2 * n3;
    const int g = ew_grid(total);
    const int passes = 1 + (n2 > 1) + (n1 > 1) + (n0 > 1);
    int* cur = (passes & 1) ? a : b;                    // ping-pong so that the last pass lands in `a`
    int* nxt = (passes & 1) ? b : a;
    if (n3 <= 256) {
        edt_first_kernel<<<g, 256, 0, st>>>(feat, n0, n1, n2, n3, bb, cur);
        SP_LAUNCH_OK("edt_first_kernel");
    } else {
        edt_pass_kernel<true><<<g, 256, 0, st>>>(feat, n0, n1, n2, n3, 3, bb, cur);
        SP_LAUNCH_OK("edt_pass_kernel<first>");
    }
    const int ext[3] = {n2, n1, n0};
    for (int ax = 0; ax < 3; ++ax) {
        if (ext[ax] == 1) continue;
        if (ext[ax] <= 256) {
            edt_tile_kernel<<<sp_num_sms() * 8, 256, (size_t)ext[ax] * 32 * sizeof(int), st>>>(cur, n0, n1, n2, n3, 2 - ax, bb, nxt);
            SP_LAUNCH_OK("edt_tile_kernel");
        } else {
            edt_pass_kernel<false><<<g, 256, 0, st>>>(cur, n0, n1, n2, n3, 2 - ax, bb, nxt);
            SP_LAUNCH_OK("edt_pass_kernel");
        }
        int* t = cur; cur = nxt; nxt = t;
    }
    return cur == a ? 0 : (sp_set_error("edt_run: internal ping-pong error"), -1);
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
    int* bb = maxd2 + 2;                                                        // [8]
    SP_CUDA(cudaMemsetAsync(counts, 0, 256, st));
    const int g = ew_grid(total);
    bbox_init_kernel<<<1, 32, 0, st>>>(bb, n0, n1, n2, n3);
    SP_LAUNCH_OK("bbox_init_kernel");

    border_kernel<<<g, 256, 0, st>>>(result, n0, n1, n2, n3, threshold, all_border, br, counts, bb);
    SP_LAUNCH_OK("border_kernel");
    border_kernel<<<g, 256, 0, st>>>(target, n0, n1, n2, n3, threshold, all_border, bt, counts + 1, bb);
    SP_LAUNCH_OK("border_kernel");
    // distances to the target's border, read at the result's border — and the other way round
    if (int e = edt_run(bt, n0, n1, n2, n3, bb, da, db, st)) return e;
    surf_reduce_kernel<<<RED_BLOCKS, 256, 0, st>>>(da, br, total, part, maxd2);
    SP_LAUNCH_OK("surf_reduce_kernel");
    if (int e = edt_run(br, n0, n1, n2, n3, bb, da, db, st)) return e;
    surf_reduce_kernel<<<RED_BLOCKS, 256, 0, st>>>(da, bt, total, part + RED_BLOCKS, maxd2 + 1);
    SP_LAUNCH_OK("surf_reduce_kernel");
    surf_final_kernel<<<1, 32, 0, st>>>(part, part + RED_BLOCKS, RED_BLOCKS, maxd2, counts, out8);
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
    if (int e = edt_run(feat, n0, n1, n2, n3, nullptr, din, tmp, st)) return e;
    // distance of background voxels to the nearest object voxel: features = object
    feature_kernel<<<g, 256, 0, st>>>(mask, total, threshold, outside_is_lt ? 2 : 1, feat);
    SP_LAUNCH_OK("feature_kernel");
    if (int e = edt_run(feat, n0, n1, n2, n3, nullptr, dout, tmp, st)) return e;
    sdm_combine_kernel<<<g, 256, 0, st>>>(din, dout, total, sign, out);
    SP_LAUNCH_OK("sdm_combine_kernel");
    return 0;
}

}  // extern "C"
