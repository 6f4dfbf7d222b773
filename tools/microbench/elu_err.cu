// elu_err.cu — rounding error of ELU(alpha=1) negative branch variants against a double reference (diagnostic).
#include <cstdio>
#include <cmath>
#include <vector>
__global__ void k(const float* v, float* a, float* b, float* c, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = expm1f(v[i]);
    b[i] = (v[i] < -1.f) ? (expf(v[i]) - 1.f) : expm1f(v[i]);
    c[i] = (float)expm1((double)v[i]);
}
int main() {
    const int n = 1 << 20;
    std::vector<float> v(n), a(n), b(n), c(n);
    for (int i = 0; i < n; ++i) v[i] = -10.f * (float)((i * 2654435761u) % 1000003u) / 1000003.f;
    float *dv, *da, *db, *dc;
    cudaMalloc(&dv, n * 4); cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&dc, n * 4);
    cudaMemcpy(dv, v.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(dv, da, db, dc, n);
    cudaMemcpy(a.data(), da, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), db, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(c.data(), dc, n * 4, cudaMemcpyDeviceToHost);
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
    return 0;
}
