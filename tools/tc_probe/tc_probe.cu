// tc_probe.cu — standalone numerics / mapping / timing probe of the tcgen05 correlation tier (sp_conv_tc.cuh).
// Diagnostic only (not part of the library): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_probe tc_probe.cu
//   ./tc_probe onehot   : one-hot weights, integer-coded input -> checks the tap / channel / voxel mapping exactly
//   ./tc_probe random   : random data vs a double-precision CPU correlation (rel-L2 per term count)
//   ./tc_probe time     : the 16->16 layer of the CAE at batch 32 x 28x124x124 -> 28x126x126 (pad 1,2,2)
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../../stroke-prediction_b200/csrc/sp_conv_tc3.cuh"

void sp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\n");
}

#define CK(x)                                                                           \
    do {                                                                                \
        cudaError_t e = (x);                                                            \
        if (e != cudaSuccess) {                                                         \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); \
            exit(2);                                                                    \
        }                                                                               \
    } while (0)

static SpConvDesc make_desc(int N, int Di, int Hi, int Wi, int Ci, int Co, int pd, int ph, int pw) {
    SpConvDesc d;
    memset(&d, 0, sizeof(d));
    d.N = N; d.Di = Di; d.Hi = Hi; d.Wi = Wi; d.Ci = Ci; d.ldi = Ci;
    d.Do = Di + 2 * pd - 2; d.Ho = Hi + 2 * ph - 2; d.Wo = Wi + 2 * pw - 2; d.Co = Co; d.ldo = Co;
    d.k = 3; d.s = 1; d.pd = pd; d.ph = ph; d.pw = pw; d.act = SP_ACT_NONE; d.alpha = 0.f;
    return d;
}

static void cpu_corr(const SpConvDesc& d, const std::vector<float>& x, const std::vector<float>& w, const std::vector<float>& sc,
                     const std::vector<float>& sh, std::vector<double>& y) {
    y.assign((size_t)d.N * d.Do * d.Ho * d.Wo * d.Co, 0.0);
    for (int n = 0; n < d.N; ++n)
        for (int od = 0; od < d.Do; ++od)
            for (int oh = 0; oh < d.Ho; ++oh)
                for (int ow = 0; ow < d.Wo; ++ow)
                    for (int co = 0; co < d.Co; ++co) {
                        double acc = 0.0;
                        for (int kd = 0; kd < 3; ++kd)
                            for (int kh = 0; kh < 3; ++kh)
                                for (int kw = 0; kw < 3; ++kw) {
                                    const int id = od - d.pd + kd, ih = oh - d.ph + kh, iw = ow - d.pw + kw;
                                    if (id < 0 || id >= d.Di || ih < 0 || ih >= d.Hi || iw < 0 || iw >= d.Wi) continue;
                                    const size_t xo = ((((size_t)n * d.Di + id) * d.Hi + ih) * d.Wi + iw) * d.Ci;
                                    for (int ci = 0; ci < d.Ci; ++ci) {
                                        float xv = x[xo + ci];
                                        if (!sc.empty()) xv = fmaf(xv, sc[ci], sh[ci]);
                                        acc += (double)xv * (double)w[((size_t)co * d.Ci + ci) * 27 + (kd * 3 + kh) * 3 + kw];
                                    }
                                }
                        y[((((size_t)n * d.Do + od) * d.Ho + oh) * d.Wo + ow) * d.Co + co] = acc;
                    }
}

static long long* g_prof = nullptr;
static int g_terms = 7;
template <int NS, int TD>
static int run_gpu(const SpConvDesc& d, const float* dx, const float* dw, const float* dsc, const float* dsh, float* dy, void* dimg) {
    if (sp_tc_pack_launch(&d, 0, NS, 16, 16, dw, dimg, 0)) return 1;
    return sp_tc_corr_launch_t<16, 16, NS, TD>(&d, d.N, dx, (const uint4*)dimg, nullptr, dsc, dsh, dy, 0, g_prof);
}

static int launch(int ns, const SpConvDesc& d, const float* dx, const float* dw, const float* dsc, const float* dsh, float* dy, void* dimg) {
    if (ns == 33 || ns == 31) {   // generation 3 (kw-stacked N, rolling depth window): three terms / bf16 mode
        const int t = ns == 33 ? 3 : 1;
        if (sp_tc3_pack_launch(&d, 0, t, 16, dw, dimg, 0)) return 1;
        return sp_tc3_corr_launch(&d, d.N, t, dx, (const uint4*)dimg, nullptr, dsc, dsh, dy, 0, g_prof);
    }
    if (ns == 32) {   // pipelined kernel (three terms, split accumulators)
        if (sp_tc_pack_launch(&d, 0, 3, 16, 16, dw, dimg, 0)) return 1;
        return sp_tc2_corr_launch(&d, d.N, dx, (const uint4*)dimg, nullptr, dsc, dsh, dy, 0, g_prof, g_terms);
    }
    return ns == 2 ? run_gpu<2, 4>(d, dx, dw, dsc, dsh, dy, dimg) : run_gpu<3, 2>(d, dx, dw, dsc, dsh, dy, dimg);
}

int main(int argc, char** argv) {
    const char* mode = argc > 1 ? argv[1] : "random";
    const int ns = argc > 2 ? atoi(argv[2]) : 2;
    if (argc > 3) g_terms = atoi(argv[3]);
    if (!strcmp(mode, "time")) {
        SpConvDesc d = make_desc(32, 28, 124, 124, 16, 16, 1, 2, 2);
        const bool full = getenv("SP_PROBE_ELU") != nullptr;        // BatchNorm prologue + ELU epilogue like the forward pass in the network
        if (full && getenv("SP_PROBE_ELU")[0] == '1') { d.act = SP_ACT_ELU; d.alpha = 1.f; }
        if (full && getenv("SP_PROBE_ELU")[0] == '3') { d.act = SP_ACT_LEAKY; d.alpha = 0.01f; }
        const size_t nx = (size_t)d.N * d.Di * d.Hi * d.Wi * d.Ci, ny = (size_t)d.N * d.Do * d.Ho * d.Wo * d.Co;
        float *dx, *dy, *dw;
        void* dimg;
        CK(cudaMalloc(&dx, nx * 4)); CK(cudaMalloc(&dy, ny * 4)); CK(cudaMalloc(&dw, 16 * 16 * 27 * 4));
        CK(cudaMalloc(&dimg, sp_tc_wimg_bytes(16, 16, 3) + sp_tc3_wimg_u4(16, 3) * 16));
        CK(cudaMemset(dx, 0, nx * 4)); CK(cudaMemset(dw, 0, 16 * 16 * 27 * 4));
        float *tsc = nullptr, *tsh = nullptr;
        if (full) {
            std::vector<float> hx(1 << 20), hw(16 * 16 * 27), one(16, 1.1f), sh(16, -0.1f);
            for (auto& v : hx) v = (float)rand() / RAND_MAX * 2.f - 1.f;
            for (auto& v : hw) v = ((float)rand() / RAND_MAX - 0.5f) * 0.3f;
            for (size_t o = 0; o < nx; o += hx.size()) CK(cudaMemcpy(dx + o, hx.data(), std::min(hx.size(), nx - o) * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMalloc(&tsc, 64)); CK(cudaMalloc(&tsh, 64));
            CK(cudaMemcpy(tsc, one.data(), 64, cudaMemcpyHostToDevice)); CK(cudaMemcpy(tsh, sh.data(), 64, cudaMemcpyHostToDevice));
        }
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int it = 0; it < 3; ++it) if (launch(ns, d, dx, dw, tsc, tsh, dy, dimg)) return 3;
        CK(cudaDeviceSynchronize());
        if (ns == 33 || ns == 31) {
            unsigned long long info[4] = {0, 0, 0, 0};
            CK(cudaMemcpyFromSymbol(info, sp_tc3::sp_tc3_dbg, sizeof(info)));
            if (info[0]) {
                printf("TIMEOUT: code 0x%llx cta %llu counter %llu parity %llu\n", info[0] & 0xffffffffull, info[0] >> 32,
                       info[1] & 0xffffffffull, info[1] >> 32);
                return 5;
            }
        }
        CK(cudaEventRecord(e0));
        const int reps = 10;
        for (int it = 0; it < reps; ++it) launch(ns, d, dx, dw, tsc, tsh, dy, dimg);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        const double flop = 2.0 * ny * 27 * 16, bytes = 4.0 * (nx + ny);
        printf("time ns=%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent)  %.0f GB/s algorithmic\n", ns, ms, flop / ms * 1e-9, bytes / ms * 1e-6);
        CK(cudaMalloc(&g_prof, 64)); CK(cudaMemset(g_prof, 0, 64));
        if (ns == 33 || ns == 31) {
            launch(ns, d, dx, dw, tsc, tsh, dy, dimg);
            CK(cudaDeviceSynchronize());
            long long hp[8];
            CK(cudaMemcpy(hp, g_prof, 64, cudaMemcpyDeviceToHost));
            const long long t = hp[3] ? hp[3] : 1;
            printf("  CTA0 cycles per output plane of issuer 0 (%lld planes = 1/3 of the CTA's): issuer: wait a_full %lld, wait t_empty %lld, issue+commit %lld | "
                   "stager (per input plane, ~3x as many): wait a_empty %lld, work %lld | epilogue group 0 (per plane, 1.5x as many): wait t_full %lld, work %lld\n",
                   t, hp[0] / t, hp[1] / t, hp[2] / t, hp[4] / (3 * t), hp[5] / (3 * t), hp[6] * 2 / (3 * t), hp[7] * 2 / (3 * t));
            return 0;
        }
        if (ns == 32) {
            launch(ns, d, dx, dw, nullptr, nullptr, dy, dimg);
            CK(cudaDeviceSynchronize());
            long long hp[8];
            CK(cudaMemcpy(hp, g_prof, 64, cudaMemcpyDeviceToHost));
            const long long t = hp[3] ? hp[3] : 1;
            printf("  CTA0 cycles per tile (%lld tiles): MMA thread: wait a_full %lld, wait t_empty %lld, issue %lld | stager: wait a_empty %lld, work %lld | "
                   "epilogue: wait t_full %lld, work %lld\n", t, hp[0] / t, hp[1] / t, hp[2] / t, hp[4] / t, hp[5] / t, hp[6] / t, hp[7] / t);
            return 0;
        }
        launch(ns, d, dx, dw, nullptr, nullptr, dy, dimg);
        CK(cudaDeviceSynchronize());
        long long hp[5];
        CK(cudaMemcpy(hp, g_prof, 40, cudaMemcpyDeviceToHost));
        printf("  CTA0 cycles per tile: stage %lld  issue %lld  mma-wait %lld  epilogue %lld  (tiles %lld)\n", hp[0] / hp[4], hp[1] / hp[4],
               hp[2] / hp[4], hp[3] / hp[4], hp[4]);
        return 0;
    }
    const bool onehot = !strcmp(mode, "onehot");
    SpConvDesc d = getenv("SP_PROBE_BIG") ? make_desc(3, 14, 45, 61, 16, 16, 1, 2, 2)
                                          : make_desc(2, 9, 37, 21, 16, 16, 1, 0, 2);     // partial tiles in every direction, padding in d and w
    const size_t nx = (size_t)d.N * d.Di * d.Hi * d.Wi * d.Ci, ny = (size_t)d.N * d.Do * d.Ho * d.Wo * d.Co;
    std::vector<float> x(nx), w((size_t)16 * 16 * 27, 0.f), sc, sh;
    srand(1234);
    if (onehot) {
        for (size_t i = 0; i < nx; ++i) x[i] = (float)((i * 7919u) % 4093u);   // integers < 2^12: exact in two bf16 terms
    } else {
        for (size_t i = 0; i < nx; ++i) x[i] = (float)rand() / RAND_MAX * 2.f - 0.7f;
        for (auto& v : w) v = ((float)rand() / RAND_MAX - 0.5f) * 0.3f;
        sc.resize(16); sh.resize(16);
        for (int c = 0; c < 16; ++c) { sc[c] = 0.5f + 0.1f * c; sh[c] = -0.3f + 0.05f * c; }
    }
    float *dx, *dy, *dw, *dsc = nullptr, *dsh = nullptr;
    void* dimg;
    CK(cudaMalloc(&dx, nx * 4)); CK(cudaMalloc(&dy, ny * 4)); CK(cudaMalloc(&dw, w.size() * 4));
    CK(cudaMalloc(&dimg, sp_tc_wimg_bytes(16, 16, 3) + sp_tc3_wimg_u4(16, 3) * 16));
    CK(cudaMemcpy(dx, x.data(), nx * 4, cudaMemcpyHostToDevice));
    if (!sc.empty()) {
        CK(cudaMalloc(&dsc, 64)); CK(cudaMalloc(&dsh, 64));
        CK(cudaMemcpy(dsc, sc.data(), 64, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dsh, sh.data(), 64, cudaMemcpyHostToDevice));
    }
    std::vector<float> y(ny);
    std::vector<double> ref;
    int bad_cases = 0;
    const int ncases = onehot ? 6 : 1;
    const int cases[6][3] = {{13, 5, 3}, {0, 0, 0}, {26, 15, 15}, {1, 8, 7}, {9, 7, 8}, {3, 12, 1}};   // tap, ci, co
    const int repeat = getenv("SP_PROBE_REPEAT") ? atoi(getenv("SP_PROBE_REPEAT")) : 1;
    if (repeat > 1) {   // stress: the same launch many times, every result compared with the first one bit for bit
        CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
        std::vector<float> y0(ny);
        int nfail = 0;
        for (int r = 0; r < repeat; ++r) {
            CK(cudaMemset(dy, 0xff, ny * 4));
            if (launch(ns, d, dx, dw, dsc, dsh, dy, dimg)) return 3;
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(y.data(), dy, ny * 4, cudaMemcpyDeviceToHost));
            if (r == 0) { y0 = y; continue; }
            size_t nd = 0, first = 0;
            for (size_t i = 0; i < ny; ++i) if (memcmp(&y[i], &y0[i], 4) != 0) { if (!nd) first = i; ++nd; }
            if (nd) {
                ++nfail;
                size_t u = first / d.Co; const int w2 = u % d.Wo; u /= d.Wo; const int h2 = u % d.Ho; u /= d.Ho;
                printf("  run %d differs from run 0 in %zu values, first at n %d od %d oh %d ow %d co %d (%g vs %g)\n", r, nd, (int)(u / d.Do),
                       (int)(u % d.Do), h2, w2, (int)(first % d.Co), y[first], y0[first]);
            }
        }
        printf("stress: %d of %d runs differ from the first\n", nfail, repeat - 1);
        return nfail ? 1 : 0;
    }
    for (int cs = 0; cs < ncases; ++cs) {
        if (onehot) {
            std::fill(w.begin(), w.end(), 0.f);
            w[((size_t)cases[cs][2] * 16 + cases[cs][1]) * 27 + cases[cs][0]] = 1.f;
        }
        CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dy, 0xff, ny * 4));
        if (launch(ns, d, dx, dw, dsc, dsh, dy, dimg)) return 3;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 4; }
        if (ns == 33 || ns == 31) {
            unsigned long long info[4] = {0, 0, 0, 0};
            CK(cudaMemcpyFromSymbol(info, sp_tc3::sp_tc3_dbg, sizeof(info)));
            if (info[0]) printf("TIMEOUT: code 0x%llx (0x1xx stager a_empty[slot], 0x2xx issuer a_full[slot + 16 iss], 0x3xx issuer t_empty[acc], "
                                "0x4xx epilogue t_full[acc + 16 grp]) cta %llu counter %llu parity %llu\n", info[0] & 0xffffffffull, info[0] >> 32,
                                info[1] & 0xffffffffull, info[1] >> 32);
        }
        CK(cudaMemcpy(y.data(), dy, ny * 4, cudaMemcpyDeviceToHost));
        cpu_corr(d, x, w, sc, sh, ref);
        double num = 0, den = 0, maxabs = 0, esum = 0, esgn = 0, esum_par[2] = {0, 0};
        size_t nbad = 0, first_bad = (size_t)-1;
        for (size_t i = 0; i < ny; ++i) {
            const double e2 = (double)y[i] - ref[i];
            esum += e2; esgn += e2 * (ref[i] > 0 ? 1.0 : -1.0);
            esum_par[(i / ((size_t)d.Co * d.Wo * d.Ho)) % d.Do & 1] += e2;
            num += e2 * e2; den += ref[i] * ref[i];
            if (fabs(e2) > maxabs) maxabs = fabs(e2);
            if (!(fabs(e2) <= 1e-3 * (1.0 + fabs(ref[i])))) { if (nbad == 0) first_bad = i; ++nbad; }
        }
        if (!onehot) {   // what an IEEE fp32 FFMA chain (the exact tier) gives on the same data
            double n32 = 0, s32 = 0, g32 = 0;
            for (int nn = 0; nn < d.N; ++nn)
                for (int od = 0; od < d.Do; ++od)
                    for (int oh = 0; oh < d.Ho; ++oh)
                        for (int ow = 0; ow < d.Wo; ++ow)
                            for (int co = 0; co < d.Co; ++co) {
                                float acc = 0.f;
                                for (int ci = 0; ci < d.Ci; ++ci)
                                    for (int kd = 0; kd < 3; ++kd)
                                        for (int kh = 0; kh < 3; ++kh)
                                            for (int kw = 0; kw < 3; ++kw) {
                                                const int id = od - d.pd + kd, ih = oh - d.ph + kh, iw = ow - d.pw + kw;
                                                if (id < 0 || id >= d.Di || ih < 0 || ih >= d.Hi || iw < 0 || iw >= d.Wi) continue;
                                                float xv = x[((((size_t)nn * d.Di + id) * d.Hi + ih) * d.Wi + iw) * d.Ci + ci];
                                                if (!sc.empty()) xv = fmaf(xv, sc[ci], sh[ci]);
                                                acc = fmaf(xv, w[((size_t)co * d.Ci + ci) * 27 + (kd * 3 + kh) * 3 + kw], acc);
                                            }
                                const double e3 = (double)acc - ref[((((size_t)nn * d.Do + od) * d.Ho + oh) * d.Wo + ow) * d.Co + co];
                                n32 += e3 * e3; s32 += e3;
                                g32 += e3 * (ref[((((size_t)nn * d.Do + od) * d.Ho + oh) * d.Wo + ow) * d.Co + co] > 0 ? 1.0 : -1.0);
                            }
            printf("fp32 FFMA chain (CPU) rel-L2 %.3e   signed error / rms: mean %.3e, mean(e * sign) %.3e\n", sqrt(n32 / (den > 0 ? den : 1)),
                   s32 / ny / sqrt(den / ny), g32 / ny / sqrt(den / ny));
        }
        if (!onehot) {
            const double rms = sqrt(den / ny);
            printf("signed error / rms(ref): mean %.3e (even planes %.3e, odd planes %.3e), mean(e * sign(ref)) %.3e   [rms error %.3e]\n",
                   esum / ny / rms, esum_par[0] / (ny / 2) / rms, esum_par[1] / (ny / 2) / rms, esgn / ny / rms, sqrt(num / ny) / rms);
        }
        if (onehot) printf("onehot tap %2d ci %2d co %2d: ", cases[cs][0], cases[cs][1], cases[cs][2]);
        printf("ns=%d rel-L2 %.3e max-abs %.3e mismatches %zu / %zu\n", ns, sqrt(num / (den > 0 ? den : 1)), maxabs, nbad, ny);
        if (nbad && !onehot) {       // where are they?  (n, od) histogram and the tiles (oh / 8, ow / 14) touched
            std::vector<int> hist((size_t)d.N * d.Do, 0);
            int shown = 0;
            for (size_t k = 0; k < ny; ++k)
                if (!(fabs((double)y[k] - ref[k]) <= 1e-3 * (1.0 + fabs(ref[k])))) {
                    size_t u = k / d.Co;
                    const int w2 = u % d.Wo; u /= d.Wo; const int h2 = u % d.Ho; u /= d.Ho;
                    ++hist[u];
                    if (shown < 12 && (k % d.Co) == 0) { printf("   bad n %d od %d oh %d ow %d (tile %d,%d; row-in-tile %d col %d) got %g exp %g\n", (int)(u / d.Do), (int)(u % d.Do), h2, w2, h2 / 8, w2 / 14, h2 % 8, w2 % 14, y[k], ref[k]); ++shown; }
                }
            for (size_t u = 0; u < hist.size(); ++u) if (hist[u]) printf("   n %d od %d: %d bad values\n", (int)(u / d.Do), (int)(u % d.Do), hist[u]);
        }
        if (nbad) {
            ++bad_cases;
            size_t i = first_bad;
            const int co = i % d.Co; size_t v = i / d.Co;
            const int ow = v % d.Wo; v /= d.Wo; const int oh = v % d.Ho; v /= d.Ho; const int od = v % d.Do; const int n = v / d.Do;
            printf("  first mismatch at n %d od %d oh %d ow %d co %d: got %g expected %g\n", n, od, oh, ow, co, y[i], ref[i]);
            if (onehot) {   // where does the value we got live in the input?
                for (size_t j = 0; j < nx; ++j)
                    if (x[j] == y[i]) {
                        const int ci = j % d.Ci; size_t u = j / d.Ci;
                        const int iw = u % d.Wi; u /= d.Wi; const int ih = u % d.Hi; u /= d.Hi; const int id = u % d.Di;
                        printf("  got-value found in input at n %d id %d ih %d iw %d ci %d\n", (int)(u / d.Di), id, ih, iw, ci);
                        break;
                    }
                int shown = 0;
                for (size_t k = 0; k < ny && shown < 8; ++k)
                    if (!(fabs((double)y[k] - ref[k]) <= 1e-3 * (1.0 + fabs(ref[k])))) {
                        const int c2 = k % d.Co; size_t u = k / d.Co;
                        const int w2 = u % d.Wo; u /= d.Wo; const int h2 = u % d.Ho; u /= d.Ho;
                        printf("   bad (n %d od %d oh %d ow %d co %d) got %g exp %g\n", (int)(u / d.Do), (int)(u % d.Do), h2, w2, c2, y[k], ref[k]);
                        ++shown;
                    }
            }
        }
    }
    return bad_cases ? 1 : 0;
}
