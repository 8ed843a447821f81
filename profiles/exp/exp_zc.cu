// Zero-copy PCIe rates from SM code on this box: a kernel reading / writing 32 MiB of page-locked host memory directly.
//   W16s : each lane stores 16 B at a 32 B stride, twice (what a naive AOS Phase store of 4 doubles does)
//   W512 : each warp instruction stores 512 contiguous bytes (lane l -> bytes [16l, 16l+16))
//   R16s / R512 : the same two shapes for loads (result folded into a checksum)
// each alone and with a copy-engine transfer running in the opposite direction.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exp_zc exp_zc.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void w16s(double2* out, long long n32) {   // n32 = number of 32-byte records
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (long long)gridDim.x * blockDim.x) {
    out[2 * i] = make_double2((double)i, 1.0);
    out[2 * i + 1] = make_double2(2.0, 3.0);
  }
}
__global__ void w512(double2* out, long long n32) {
  const long long n16 = 2 * n32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    out[i] = make_double2((double)i, 1.0);
}
__global__ void r16s(const double2* in, long long n32, double* sink) {
  double acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (long long)gridDim.x * blockDim.x) {
    double2 a = in[2 * i], b = in[2 * i + 1];
    acc += a.x + a.y + b.x + b.y;
  }
  if (acc == 12345.678) *sink = acc;
}
__global__ void r512(const double2* in, long long n32, double* sink) {
  const long long n16 = 2 * n32;
  double acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    double2 a = in[i];
    acc += a.x + a.y;
  }
  if (acc == 12345.678) *sink = acc;
}
int main() {
  const size_t bytes = 32u << 20;
  const long long n32 = bytes / 32;
  void *h_a, *h_b, *d_a, *d_b; double* sink;
  CK(cudaHostAlloc(&h_a, bytes, cudaHostAllocDefault));
  CK(cudaHostAlloc(&h_b, bytes, cudaHostAllocDefault));
  CK(cudaMalloc(&d_a, bytes)); CK(cudaMalloc(&d_b, bytes)); CK(cudaMalloc(&sink, 8));
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int grid : {148, 888, 3552}) for (int variant = 0; variant < 4; variant++) for (int duplex = 0; duplex < 2; duplex++) {
    float best = 1e9;
    for (int rep = 0; rep < 6; rep++) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, s1));
      if (duplex) { if (variant < 2) CK(cudaMemcpyAsync(d_b, h_b, bytes, cudaMemcpyHostToDevice, s2)); else CK(cudaMemcpyAsync(h_b, d_b, bytes, cudaMemcpyDeviceToHost, s2)); }
      if (variant == 0) w16s<<<grid, 128, 0, s1>>>((double2*)h_a, n32);
      if (variant == 1) w512<<<grid, 128, 0, s1>>>((double2*)h_a, n32);
      if (variant == 2) r16s<<<grid, 128, 0, s1>>>((const double2*)h_a, n32, sink);
      if (variant == 3) r512<<<grid, 128, 0, s1>>>((const double2*)h_a, n32, sink);
      CK(cudaEventRecord(e1, s1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep && ms < best) best = ms;
    }
    const char* nm[4] = {"W16s", "W512", "R16s", "R512"};
    printf("grid %4d %s %s: %.3f ms -> %.1f GB/s\n", grid, nm[variant], duplex ? "with CE copy the other way" : "alone                     ", best, bytes / best / 1e6);
  }
  return 0;
}
