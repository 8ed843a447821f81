// Issue model of the FP64 pipe next to other instructions (round 2): do integer / shared-memory / FP32 instructions issue
// "for free" in the clocks a DFMA keeps the FP64 pipe busy, or does every instruction cost its own issue clock?
//   DF = 0: x = fma(x, y, const)  (2 register operands, 2 clk)     DF = 1: x = fma(x, y, z)  (3 register operands, 3 clk)
//   per group of 8 independent DFMAs, NX extra instructions of kind KX are mixed in:
//   KX = 0: integer LOP3/IADD3 on 4 independent chains   1: LDS.64 (conflict-free)   2: FFMA   3: MUFU.RCP (fp32)   4: IMAD
// Reports clocks per 8-DFMA group per scheduler; run with 8 and with 5 warps per scheduler (the RK4 kernel has 5).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exp_fp64c exp_fp64c.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int DF, int KX, int NX>
__global__ void k(double* out, long long* cyc, const double* in, int iters, double b) {
  constexpr int CH = 8;
  __shared__ double sm[2048];
  double x[CH], y[CH], z[CH];
  unsigned m[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
  float f[4] = {1.0f + threadIdx.x, 2.0f, 3.0f, 4.0f};
  double acc = 0.0;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
#pragma unroll
  for (int c = 0; c < CH; c++) { x[c] = in[threadIdx.x + 32 * c]; y[c] = in[threadIdx.x + 32 * c + 1024]; z[c] = in[threadIdx.x + 32 * c + 2048]; }
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int c = 0; c < CH; c++) {
        if (DF == 0) x[c] = fma(x[c], y[c], b);
        else x[c] = fma(x[c], y[c], z[(c + u) % CH]);
      }
#pragma unroll
      for (int q = 0; q < NX; q++) {
        if (KX == 0) m[q & 3] = (m[q & 3] ^ (m[(q + 1) & 3] | 0x55u)) + u;
        if (KX == 1) acc += sm[(threadIdx.x + 32 * ((q + u + m[0]) & 31)) & 2047];
        if (KX == 2) f[q & 3] = fmaf(f[q & 3], 1.0001f, 0.5f);
        if (KX == 3) f[q & 3] = __frcp_rn(f[q & 3]) + 1.0f;
        if (KX == 4) m[q & 3] = m[q & 3] * 1664525u + 1013904223u;
      }
    }
  }
  const long long t1 = clock64();
  double s = (double)(m[0] + m[1] + m[2] + m[3]) + f[0] + f[1] + f[2] + f[3] + acc;
#pragma unroll
  for (int c = 0; c < CH; c++) s += x[c];
  if (s == 1.2345) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int DF, int KX, int NX>
void run(int sms, const double* in, int block) {
  const int grid = sms, iters = 1024;
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * grid);
  k<DF, KX, NX><<<grid, block>>>(out, cyc, in, 16, 1e-9);
  k<DF, KX, NX><<<grid, block>>>(out, cyc, in, iters, 1e-9);
  cudaDeviceSynchronize();
  long long* h = new long long[grid]; cudaMemcpy(h, cyc, 8 * grid, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; i++) avg += h[i]; avg /= grid;
  const double warps_per_sched = block / 128.0;
  // clocks one scheduler spends per (8 DFMA + NX extra) group of one warp
  const double clk = avg / ((double)iters * 8 * warps_per_sched);
  static const char* kn[] = {"LOP3+IADD", "LDS.64", "FFMA", "MUFU+FADD", "IMAD"};
  printf("warps/sched %.0f  DFMA %d-reg  + %2d x %-9s per 8 DFMA: %6.2f clk per group  (DFMA alone would be %d)\n", warps_per_sched,
         DF ? 3 : 2, NX * (KX == 0 ? 2 : KX == 3 ? 2 : 1), kn[KX], clk, DF ? 24 : 16);
  cudaFree(out); cudaFree(cyc); delete[] h;
}
template <int DF, int KX> void sweep(int sms, const double* in, int block) {
  run<DF, KX, 0>(sms, in, block); run<DF, KX, 2>(sms, in, block); run<DF, KX, 4>(sms, in, block); run<DF, KX, 8>(sms, in, block);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* in; cudaMalloc(&in, 8 * 4096);
  double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0 + 1e-9 * (i % 97);
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  for (int block : {1024, 640}) {
    sweep<0, 0>(sms, in, block); sweep<1, 0>(sms, in, block);
    sweep<0, 1>(sms, in, block); sweep<0, 2>(sms, in, block); sweep<0, 3>(sms, in, block); sweep<0, 4>(sms, in, block);
  }
  return 0;
}
