// What the FP64 pipe of this B200 really sustains: DFMA lanes per clock per SM (from clock64 inside the kernel, so the
// figure is independent of the SM clock) and per second, for 1..8 warps per scheduler and 1..8 independent chains per thread.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exp_fp64 exp_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 1e-3 + c;
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int c = 0; c < CH; c++) x[c] = fma(x[c], a, b);
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += x[c];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
void run(int warps_per_smsp, int sms) {
  const int block = 128 * warps_per_smsp > 1024 ? 1024 : 128 * warps_per_smsp;     // 4 SMSPs per SM
  const int ctas_per_sm = (128 * warps_per_smsp + block - 1) / block;
  const int grid = sms * ctas_per_sm, iters = 4096;
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * grid);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<grid, block>>>(out, cyc, 64, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<CH><<<grid, block>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[grid]; cudaMemcpy(h, cyc, 8 * grid, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
  const double lanes = (double)iters * 8 * CH * block * ctas_per_sm;      // DFMA lane-ops per SM
  printf("chains %d  warps/SMSP %d : %.1f DFMA lanes/clk/SM (clock64), %.2f TFLOP/s (events), implied clock %.0f MHz\n", CH, warps_per_smsp,
         lanes / avg, 2.0 * lanes * sms / (ms * 1e-3) / 1e12, avg / (ms * 1e-3) / 1e6);
  cudaFree(out); cudaFree(cyc); delete[] h;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  for (int w : {1, 2, 4, 6, 8}) { run<1>(w, sms); run<2>(w, sms); run<4>(w, sms); run<8>(w, sms); }
  return 0;
}
