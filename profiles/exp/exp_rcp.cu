// Accuracy experiment: MUFU.RCP64H seed + (a) two Newton steps (4 FMA) vs (b) one cubic step (3 FMA) vs IEEE 1/d.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_rcp exp_rcp.cu ; prints max relative error in ulps.
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ double seed(double d) { double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d)); return x; }
__device__ __forceinline__ double rcp_a(double d) { double x = seed(d); double e = fma(-d, x, 1.0); x = fma(x, e, x); e = fma(-d, x, 1.0); return fma(x, e, x); }
__device__ __forceinline__ double rcp_b(double d) { double x = seed(d); double e = fma(-d, x, 1.0); double t = fma(e, e, e); return fma(x, t, x); }
__device__ unsigned long long g_max[4];
__global__ void k(unsigned long long n) {
  unsigned long long z = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 12345;
  double ma = 0, mb = 0, ms = 0;
  for (unsigned long long it = 0; it < n; it++) {
    z += 0x9E3779B97F4A7C15ULL; unsigned long long x = z; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; x ^= x >> 31;
    // mantissa uniformly random, exponent in a +-40 window around 1
    long long ex = 1023 + (long long)((x >> 52) % 81) - 40;
    double d = __longlong_as_double((x & 0x800fffffffffffffULL) | ((unsigned long long)ex << 52));
    double r = 1.0 / d;
    double ulp = fabs(r) * 1.1102230246251565e-16;
    ma = fmax(ma, fabs(rcp_a(d) - r) / ulp); mb = fmax(mb, fabs(rcp_b(d) - r) / ulp); ms = fmax(ms, fabs(seed(d) - r) / fabs(r));
  }
  atomicMax(&g_max[0], (unsigned long long)(ma * 1000)); atomicMax(&g_max[1], (unsigned long long)(mb * 1000));
  atomicMax(&g_max[2], (unsigned long long)(ms * 1e12));
}
int main() {
  k<<<148 * 8, 256>>>(20000);
  unsigned long long h[4];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_max, sizeof h);
  printf("samples %.3g | max err (half-ulps of result): newton2 %.3f  cubic1 %.3f | seed max rel err %.3e (2^%.1f)\n", 148.0 * 8 * 256 * 20000, h[0] / 1000.0, h[1] / 1000.0, h[2] / 1e12, log2(h[2] / 1e12));
  return 0;
}
