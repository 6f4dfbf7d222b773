// elu_err.cu — rounding error of ELU(alpha=1) negative branch variants against a double reference (diagnostic).
#include <cstdio>
#include <cmath>
#include <vector>
#include "../../stroke-prediction_b200/csrc/sp_common.cuh"
__global__ void k(const float* v, float* a, float* b, float* c, float* e, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = expm1f(v[i]);
    b[i] = (v[i] < -1.f) ? (expf(v[i]) - 1.f) : expm1f(v[i]);
    c[i] = (float)expm1((double)v[i]);
    e[i] = sp_expm1_neg(v[i]);          // the product path's ELU branch (sp_common.cuh)
}
int main() {
    const int n = 1 << 20;
    std::vector<float> v(n), a(n), b(n), c(n), e(n);
    for (int i = 0; i < n; ++i) v[i] = -10.f * (float)((i * 2654435761u) % 1000003u) / 1000003.f;
    float *dv, *da, *db, *dc, *de;
    cudaMalloc(&dv, n * 4); cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&dc, n * 4); cudaMalloc(&de, n * 4);
    cudaMemcpy(dv, v.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(dv, da, db, dc, de, n);
    cudaMemcpy(a.data(), da, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), db, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(c.data(), dc, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(e.data(), de, n * 4, cudaMemcpyDeviceToHost);
    for (int band = 0; band < 5; ++band) {
        double ea = 0, eb = 0, ec = 0, ecpu = 0; int m = 0;
        for (int i = 0; i < n; ++i) {
            if ((int)(-v[i] / 2.f) != band) continue;
            const double r = expm1((double)v[i]);
            const double ulp = ldexp(1.0, ilogb(r) - 23);
            ea += pow((a[i] - r) / ulp, 2); eb += pow((b[i] - r) / ulp, 2); ec += pow((c[i] - r) / ulp, 2);
            ecpu += pow((expm1f(v[i]) - r) / ulp, 2); ++m;
        }
        printf("v in [-%d,-%d): rms ulp error  expm1f(gpu) %.3f  hybrid %.3f  double->float %.3f  expm1f(host libm) %.3f\n", 2 * band + 2, 2 * band,
               sqrt(ea / m), sqrt(eb / m), sqrt(ec / m), sqrt(ecpu / m));
    }
    double worst = 0, worst_lib = 0, rms = 0; float at = 0;
    for (int i = 0; i < n; ++i) {
        if (v[i] == 0.f) continue;
        const double r = expm1((double)v[i]);
        const double re = fabs((e[i] - r) / r), rl = fabs((a[i] - r) / r);
        rms += re * re;
        if (re > worst) { worst = re; at = v[i]; }
        if (rl > worst_lib) worst_lib = rl;
    }
    printf("sp_expm1_neg on (-10, 0): max relative error %.3e at v = %.4f, rms %.3e   (expm1f on the device: max %.3e)\n", worst, at, sqrt(rms / n), worst_lib);
    return 0;
}
