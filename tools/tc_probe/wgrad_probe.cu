// wgrad_probe.cu — standalone mapping / numerics / timing probe of the tcgen05 weight-gradient tier (sp_wgrad_tc.cuh).
// Diagnostic only: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o wgrad_probe wgrad_probe.cu
//   ./wgrad_probe check [drain_every]  : small ragged geometry (padding in d / w, partial tiles) vs a double-precision CPU sum
//   modes: check | time (16 channels; check4 / time4 = second generation, sp_wgrad_tc4.cuh), check24 | time24 (24 channels); suffix 'a' (checka, time24a) = experimental A-in-TMEM kernel
//   ./wgrad_probe time  [drain_every]  : the 16->16 layer of the CAE decoder at batch 32 (28x126x126 -> 28x128x128, pad 1,2,2)
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "sp_wgrad_tca.cuh"
#include "../../stroke-prediction_b200/csrc/sp_wgrad_tc4.cuh"
#include "../../stroke-prediction_b200/csrc/sp_wgrad_tc4s2.cuh"

void sp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fprintf(stderr, "\n");
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int chunks, int64_t wn, float* __restrict__ dw, float beta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wn) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += ws[(int64_t)c * wn + i];
    dw[i] = (beta != 0.f ? beta * dw[i] : 0.f) + s;
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

static bool g_wide = false, g_tca = false, g_tc4 = false, g_s2 = false;
static int launch_wgrad(const SpConvDesc* d, const float* dx, const float* dsc, const float* dsh, const float* dz, float* ddw, float* dws,
                        long long* prof, int drain_every, int mrows) {
    if (g_s2) return sp_tc4s2_wgrad_launch(d, d->N, dx, dsc, dsh, dz, nullptr, nullptr, ddw, 0.f, dws, 0, prof, drain_every);
    if (g_tc4) return sp_tc4_wgrad_launch(d, d->N, dx, dsc, dsh, dz, nullptr, nullptr, ddw, 0.f, dws, 0, prof, drain_every);
    if (g_tca) return sp_tca_wgrad_launch(d, d->N, dx, dsc, dsh, dz, nullptr, nullptr, ddw, 0.f, dws, 0, prof, drain_every);
    if (g_wide) return sp_tc24_wgrad_launch(d, d->N, dx, dsc, dsh, dz, nullptr, nullptr, ddw, 0.f, dws, 0, prof, drain_every);
    return sp_tc_wgrad_launch(d, d->N, dx, dsc, dsh, dz, nullptr, nullptr, ddw, 0.f, dws, 0, prof, drain_every, mrows);
}

int main(int argc, char** argv) {
    char modebuf[32];
    strncpy(modebuf, argc > 1 ? argv[1] : "check", 31); modebuf[31] = 0;
    if (strlen(modebuf) > 1 && modebuf[strlen(modebuf) - 1] == 'a') { g_tca = true; modebuf[strlen(modebuf) - 1] = 0; }   // checka / time24a: A in TMEM
    if (strlen(modebuf) > 2 && !strcmp(modebuf + strlen(modebuf) - 2, "s2")) { g_s2 = true; modebuf[strlen(modebuf) - 2] = 0; }   // checks2 / times2: stride-2 layer
    if (strlen(modebuf) > 1 && modebuf[strlen(modebuf) - 1] == '4' && modebuf[strlen(modebuf) - 2] != '2') { g_tc4 = true; modebuf[strlen(modebuf) - 1] = 0; }   // check4 / time4: second generation
    const char* mode = modebuf;
    const int drain_every = argc > 2 ? atoi(argv[2]) : 2;
    const int mrows = argc > 3 ? atoi(argv[3]) : 128;
    const bool timing = !strncmp(mode, "time", 4);
    const bool wide = strlen(mode) > 2 && !strcmp(mode + strlen(mode) - 2, "24");      // check24 / time24: the 24-channel tier
    SpConvDesc d = wide ? (timing ? make_desc(32, 14, 58, 58, 24, 24, 1, 2, 2) : make_desc(2, 9, 21, 45, 24, 20, 1, 0, 2))
                        : (timing ? make_desc(32, 28, 126, 126, 16, 16, 1, 2, 2) : make_desc(2, 9, 21, 45, 16, 16, 1, 0, 2));
    g_wide = wide;
    if (g_s2) {      // Cae3D.py:48: 16 -> 24, k3 s2 p1 (check: odd depth, ragged tiles)
        d = timing ? make_desc(24, 28, 124, 124, 16, 24, 1, 1, 1) : make_desc(2, 17, 22, 90, 16, 24, 1, 1, 1);
        d.s = 2;
        d.Do = (d.Di + 2 - 3) / 2 + 1; d.Ho = (d.Hi + 2 - 3) / 2 + 1; d.Wo = (d.Wi + 2 - 3) / 2 + 1;
    }
    const size_t nx = (size_t)d.N * d.Di * d.Hi * d.Wi * d.Ci, nz = (size_t)d.N * d.Do * d.Ho * d.Wo * d.Co;
    const int wn = d.Co * d.Ci * 27;
    const sp_wtc::WtcPlan p = sp_wtc::plan(&d);
    printf("geometry N %d I %dx%dx%d O %dx%dx%d pad %d,%d,%d  tiles %lld grid %d smem %zu\n", d.N, d.Di, d.Hi, d.Wi, d.Do, d.Ho, d.Wo,
           d.pd, d.ph, d.pw, (long long)p.total, p.grid, (size_t)sp_wtc::SMEM_W);
    std::vector<float> x(nx), z(nz), sc(24), sh(24);
    srand(4321);
    for (size_t i = 0; i < nx; ++i) x[i] = (float)rand() / RAND_MAX * 2.f - 0.7f;
    for (size_t i = 0; i < nz; ++i) z[i] = ((float)rand() / RAND_MAX - 0.45f) * 0.01f;
    for (int c = 0; c < 24; ++c) { sc[c] = 0.5f + 0.1f * c; sh[c] = -0.3f + 0.05f * c; }
    float *dx, *dz, *dsc, *dsh, *dws, *ddw;
    long long* dprof;
    CK(cudaMalloc(&dx, nx * 4)); CK(cudaMalloc(&dz, nz * 4)); CK(cudaMalloc(&dsc, 96)); CK(cudaMalloc(&dsh, 96));
    CK(cudaMalloc(&dws, (size_t)148 * wn * 4 + (size_t)148 * 16 * 6912 * 4)); CK(cudaMalloc(&ddw, wn * 4)); CK(cudaMalloc(&dprof, 128));
    CK(cudaMemcpy(dx, x.data(), nx * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dz, z.data(), nz * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dsc, sc.data(), 96, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dsh, sh.data(), 96, cudaMemcpyHostToDevice));
    CK(cudaMemset(dprof, 0, 128));

    if (timing) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int it = 0; it < 2; ++it)
            if (launch_wgrad(&d, dx, dsc, dsh, dz, ddw, dws, nullptr, drain_every, mrows)) return 3;
        CK(cudaDeviceSynchronize());
        const int reps = 5;
        CK(cudaEventRecord(e0));
        for (int it = 0; it < reps; ++it) launch_wgrad(&d, dx, dsc, dsh, dz, ddw, dws, nullptr, drain_every, mrows);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        const double flop = 2.0 * nz * 27 * d.Ci, bytes = 4.0 * (nx + nz);
        printf("time drain_every=%d mrows=%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent)  %.0f GB/s algorithmic\n", drain_every, mrows, ms, flop / ms * 1e-9,
               bytes / ms * 1e-6);
        launch_wgrad(&d, dx, dsc, dsh, dz, ddw, dws, dprof, drain_every, mrows);
        CK(cudaDeviceSynchronize());
        long long hp[16];
        CK(cudaMemcpy(hp, dprof, 128, cudaMemcpyDeviceToHost));
        const long long t = hp[3] ? hp[3] : 1;
        printf("  CTA0 cycles per step (%lld steps): issuer 0: wait a_full %lld, wait t_empty %lld, issue %lld | stager: wait a_empty %lld, work %lld | "
               "drain warp 0: wait t_full %lld, work %lld | issuer s_full wait %lld | producer: wait s_free %lld, work %lld\n", t, hp[0] / t, hp[1] / t, hp[2] / t, hp[4] / t, hp[5] / t, hp[6] / t, hp[7] / t, hp[8] / t, hp[9] / t, hp[10] / t);
        printf("  CTA0 time line (cycles after the prologue): issuer loop done %lld, drain + flush done %lld, exit %lld\n", hp[10], hp[11], hp[12]);
        return 0;
    }

    if (launch_wgrad(&d, dx, dsc, dsh, dz, ddw, dws, nullptr, drain_every, mrows)) return 3;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 4; }
    std::vector<float> dw(wn);
    CK(cudaMemcpy(dw.data(), ddw, wn * 4, cudaMemcpyDeviceToHost));
    // CPU: double reference and a sequential fp32 FFMA chain
    std::vector<double> ref(wn, 0.0);
    std::vector<float> f32(wn, 0.f);
    for (int n = 0; n < d.N; ++n)
        for (int od = 0; od < d.Do; ++od)
            for (int oh = 0; oh < d.Ho; ++oh)
                for (int ow = 0; ow < d.Wo; ++ow) {
                    const float* zp = &z[((((size_t)n * d.Do + od) * d.Ho + oh) * d.Wo + ow) * d.Co];
                    for (int kd = 0; kd < 3; ++kd)
                        for (int kh = 0; kh < 3; ++kh)
                            for (int kw = 0; kw < 3; ++kw) {
                                const int id = od * d.s - d.pd + kd, ih = oh * d.s - d.ph + kh, iw = ow * d.s - d.pw + kw;
                                if (id < 0 || id >= d.Di || ih < 0 || ih >= d.Hi || iw < 0 || iw >= d.Wi) continue;
                                const float* xp = &x[((((size_t)n * d.Di + id) * d.Hi + ih) * d.Wi + iw) * d.Ci];
                                const int tap = (kd * 3 + kh) * 3 + kw;
                                for (int ci = 0; ci < d.Ci; ++ci) {
                                    const float xv = fmaf(xp[ci], sc[ci], sh[ci]);
                                    for (int co = 0; co < d.Co; ++co) {
                                        const int i = (co * d.Ci + ci) * 27 + tap;
                                        ref[i] += (double)xv * (double)zp[co];
                                        f32[i] = fmaf(xv, zp[co], f32[i]);
                                    }
                                }
                            }
                }
    double num = 0, den = 0, n32 = 0, maxabs = 0;
    int nbad = 0, first_bad = -1;
    double scale_ref = 0;
    for (int i = 0; i < wn; ++i) scale_ref = fmax(scale_ref, fabs(ref[i]));
    for (int i = 0; i < wn; ++i) {
        const double e2 = (double)dw[i] - ref[i], e3 = (double)f32[i] - ref[i];
        num += e2 * e2; den += ref[i] * ref[i]; n32 += e3 * e3;
        maxabs = fmax(maxabs, fabs(e2));
        if (!(fabs(e2) <= 1e-4 * scale_ref)) { if (nbad == 0) first_bad = i; ++nbad; }
    }
    printf("tc wgrad rel-L2 %.3e (fp32 FFMA chain on the CPU: %.3e)  max-abs %.3e (max |ref| %.3e)  mismatches %d / %d\n",
           sqrt(num / den), sqrt(n32 / den), maxabs, scale_ref, nbad, wn);
    if (nbad) {
        int shown = 0;
        for (int i = first_bad; i < wn && shown < 12; ++i)
            if (!(fabs((double)dw[i] - ref[i]) <= 1e-4 * scale_ref)) {
                printf("   bad co %d ci %d tap %d (kd %d kh %d kw %d): got %g expected %g\n", i / (27 * d.Ci), (i / 27) % d.Ci, i % 27,
                       (i % 27) / 9, ((i % 27) / 3) % 3, i % 3, dw[i], ref[i]);
                ++shown;
            }
    }
    return nbad ? 1 : 0;
}
