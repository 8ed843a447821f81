// A/B for DESIGN.md section 3 / VERDICT r1 item 3: thread-per-trajectory vs sub-warp cooperative (LANES lanes per trajectory,
// warp shuffles) for the dense part of the (24, 12) chain's hamEqs: assemble the 12 x 12 mass matrix
// M_ij = (12 - max(i, j)) cos(th_i - th_j) from the angles, factor it (LDL^T), solve M v = p.
//   T        one thread owns one system: packed lower triangle (78 doubles) in registers, fully unrolled LDL^T
//   L<LANES> LANES lanes own one system: lane l holds rows i = l, l + LANES, ...; the pivot column travels by __shfl_sync
//            (width = LANES), every lane updates its own rows, substitutions broadcast one unknown per step
// Both variants run REPS dependent solves per system (th <- th + 1e-3 v) and are checked against a host long-double solve.
// Output: systems x solves per second and the largest residual.  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cmath>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

constexpr int N = 12;
template <int I> struct Idx { static constexpr int value = I; };
template <int B, int E, class F> __device__ __forceinline__ void sfor(F&& f) {
  if constexpr (E - B == 1) f(Idx<B>{});
  else if constexpr (E - B > 1) { sfor<B, (B + E) / 2>(f); sfor<(B + E) / 2, E>(f); }
}
#define IDX(n, t) constexpr int n = decltype(t)::value
__device__ __forceinline__ constexpr int tri(int j, int k) { return j * (j + 1) / 2 + k; }
__device__ __forceinline__ double rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  const double e = fma(-d, x, 1.0);
  return fma(x, fma(e, e, e), x);
}

// ---- T: one thread per system ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) solve_thread(const double* __restrict__ th0, const double* __restrict__ p0, double* __restrict__ out, int nsys, int reps) {
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= nsys) return;
  double th[N], p[N], v[N];
#pragma unroll
  for (int j = 0; j < N; j++) { th[j] = th0[(size_t)i0 * N + j]; p[j] = p0[(size_t)i0 * N + j]; }
  for (int r = 0; r < reps; r++) {
    double s[N], c[N], A[N * (N + 1) / 2];
#pragma unroll
    for (int j = 0; j < N; j++) sincos(th[j], &s[j], &c[j]);
    sfor<0, N>([&](auto jt) { IDX(j, jt); sfor<0, j + 1>([&](auto kt) { IDX(k, kt); A[tri(j, k)] = (double)(N - j) * fma(c[j], c[k], s[j] * s[k]); }); });
    double invd[N];
    sfor<0, N>([&](auto jt) {
      IDX(j, jt);
      double w[j > 0 ? j : 1];
      double d = A[tri(j, j)];
      sfor<0, j>([&](auto kt) { IDX(k, kt); w[k] = A[tri(j, k)] * A[tri(k, k)]; d = fma(-A[tri(j, k)], w[k], d); });
      A[tri(j, j)] = d;
      const double id = rcp(d);
      invd[j] = id;
      sfor<j + 1, N>([&](auto it) {
        IDX(i, it);
        double t = A[tri(i, j)];
        sfor<0, j>([&](auto kt) { IDX(k, kt); t = fma(-A[tri(i, k)], w[k], t); });
        A[tri(i, j)] = t * id;
      });
    });
    sfor<0, N>([&](auto jt) { IDX(j, jt); double t = p[j]; sfor<0, j>([&](auto kt) { IDX(k, kt); t = fma(-A[tri(j, k)], v[k], t); }); v[j] = t; });
#pragma unroll
    for (int j = 0; j < N; j++) v[j] *= invd[j];
    sfor<0, N>([&](auto rt) { IDX(rr, rt); constexpr int j = N - 1 - rr; double t = v[j]; sfor<j + 1, N>([&](auto kt) { IDX(k, kt); t = fma(-A[tri(k, j)], v[k], t); }); v[j] = t; });
    if (r + 1 < reps) {
#pragma unroll
      for (int j = 0; j < N; j++) th[j] = fma(1e-3, v[j], th[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < N; j++) out[(size_t)i0 * N + j] = v[j];
}

// ---- L: LANES lanes per system, rows distributed cyclically, shuffles --------------------------------------------------
__device__ __forceinline__ double shfl(double x, int src, int width) { return __shfl_sync(0xffffffffu, x, src, width); }
template <int LANES>
__global__ void __launch_bounds__(128) solve_lanes(const double* __restrict__ th0, const double* __restrict__ p0, double* __restrict__ out, int nsys, int reps) {
  constexpr int NR = (N + LANES - 1) / LANES;                 // row slots per lane
  const int l = threadIdx.x % LANES;
  const int sys = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  const bool live = sys < nsys;
  const int sy = live ? sys : 0;
  // row slot c holds row i = l + LANES c (all 12 columns kept; only k <= i is meaningful)
  double th[NR], p[NR], row[NR][N], v[NR];
#pragma unroll
  for (int c = 0; c < NR; c++) { const int i = l + LANES * c; th[c] = i < N ? th0[(size_t)sy * N + i] : 0.0; p[c] = i < N ? p0[(size_t)sy * N + i] : 0.0; }
  for (int r = 0; r < reps; r++) {
    double s[NR], cs[NR];
#pragma unroll
    for (int c = 0; c < NR; c++) sincos(th[c], &s[c], &cs[c]);
    // every lane needs sin/cos of all angles k <= its rows: all-gather by shuffles (source lane and slot are static)
    double sa[N], ca[N];
    sfor<0, N>([&](auto kt) { IDX(k, kt); sa[k] = shfl(s[k / LANES], k % LANES, LANES); ca[k] = shfl(cs[k / LANES], k % LANES, LANES); });
    sfor<0, NR>([&](auto ct) {
      IDX(c, ct);
      const int i = l + LANES * c;
      sfor<0, N>([&](auto kt) { IDX(k, kt); row[c][k] = (double)(N - (i > k ? i : k)) * fma(cs[c], ca[k], s[c] * sa[k]); });
    });
    // right-looking LDL^T: step k broadcasts column k (rows k..11, unscaled) and every lane updates its own rows
    double invd[N];
    sfor<0, N>([&](auto kt) {
      IDX(k, kt);
      double col[N];
      sfor<k, N>([&](auto jt) { IDX(j, jt); col[j] = shfl(row[j / LANES][k], j % LANES, LANES); });
      const double id = rcp(col[k]);
      invd[k] = id;
      sfor<0, NR>([&](auto ct) {
        IDX(c, ct);
        const int i = l + LANES * c;
        const double lik = row[c][k] * id;
        sfor<k + 1, N>([&](auto jt) { IDX(j, jt); if (j <= i) row[c][j] = fma(-lik, col[j], row[c][j]); });
        if (i > k) row[c][k] = lik;
      });
    });
    // forward substitution L y = p: the owner of row k finishes y_k and broadcasts it
    double y[NR];
#pragma unroll
    for (int c = 0; c < NR; c++) y[c] = p[c];
    sfor<0, N>([&](auto kt) {
      IDX(k, kt);
      const double yk = shfl(y[k / LANES], k % LANES, LANES);
      sfor<0, NR>([&](auto ct) { IDX(c, ct); if (l + LANES * c > k) y[c] = fma(-row[c][k], yk, y[c]); });
    });
    sfor<0, NR>([&](auto ct) { IDX(c, ct); const int i = l + LANES * c; double dsel = 0.0; sfor<0, N>([&](auto kt) { IDX(k, kt); if (k == i) dsel = invd[k]; }); y[c] *= dsel; });
    // backward substitution L^T x = z: x_k known -> row k's owner sends l_ki x_k to the owner of every row i < k
#pragma unroll
    for (int c = 0; c < NR; c++) v[c] = y[c];
    sfor<0, N>([&](auto rt) {
      IDX(rr, rt);
      constexpr int k = N - 1 - rr;
      const double xk = shfl(v[k / LANES], k % LANES, LANES);
      sfor<0, k>([&](auto it) {
        IDX(i, it);
        const double t = shfl(row[k / LANES][i] * xk, k % LANES, LANES);     // l_ki x_k from row k's owner
        if (l == i % LANES) v[i / LANES] -= t;
      });
    });
    if (r + 1 < reps) {
#pragma unroll
      for (int c = 0; c < NR; c++) th[c] = fma(1e-3, v[c], th[c]);
    }
  }
  if (live) {
#pragma unroll
    for (int c = 0; c < NR; c++) { const int i = l + LANES * c; if (i < N) out[(size_t)sys * N + i] = v[c]; }
  }
}

static void host_solve(const double* th, const double* p, int reps, double* v) {
  long double t[N], A[N][N], b[N], x[N];
  for (int j = 0; j < N; j++) t[j] = th[j];
  for (int r = 0; r < reps; r++) {
    for (int i = 0; i < N; i++) { b[i] = p[i]; for (int j = 0; j < N; j++) A[i][j] = (long double)(N - (i > j ? i : j)) * cosl(t[i] - t[j]); }
    for (int k = 0; k < N; k++) for (int i = k + 1; i < N; i++) { const long double f = A[i][k] / A[k][k]; for (int j = k; j < N; j++) A[i][j] -= f * A[k][j]; b[i] -= f * b[k]; }
    for (int i = N - 1; i >= 0; i--) { long double s = b[i]; for (int j = i + 1; j < N; j++) s -= A[i][j] * x[j]; x[i] = s / A[i][i]; }
    if (r + 1 < reps) for (int j = 0; j < N; j++) t[j] += 1e-3L * x[j];
  }
  for (int j = 0; j < N; j++) v[j] = (double)x[j];
}

template <class K>
static void run(const char* name, K kern, int lanes, const double* th, const double* p, double* out, int nsys, int reps, const std::vector<double>& hth, const std::vector<double>& hp) {
  const int threads = 128;
  const long long total = (long long)nsys * lanes;
  const int blocks = (int)((total + threads - 1) / threads);
  kern<<<blocks, threads>>>(th, p, out, nsys, reps);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) kern<<<blocks, threads>>>(th, p, out, nsys, reps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  std::vector<double> hv((size_t)64 * N);
  cudaMemcpy(hv.data(), out, sizeof(double) * 64 * N, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int s = 0; s < 64; s++) { double ref[N]; host_solve(&hth[(size_t)s * N], &hp[(size_t)s * N], reps, ref); for (int j = 0; j < N; j++) worst = fmax(worst, fabs(ref[j] - hv[(size_t)s * N + j])); }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  printf("%-28s %8.3f ms  %.3e assemble+factor+solve per s   %3d registers, %5zu B local   max |v - v_ref| %.2e   (%s)\n", name, ms,
         (double)nsys * reps / (ms * 1e-3), fa.numRegs, fa.localSizeBytes, worst, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int nsys = 1 << 18, reps = 8;
  std::vector<double> hth((size_t)nsys * N), hp((size_t)nsys * N);
  unsigned long long st = 0x9E3779B97F4A7C15ULL;
  auto rnd = [&] { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) * (1.0 / 9007199254740992.0); };
  for (auto& x : hth) x = (2 * rnd() - 1) * 3.14159;
  for (auto& x : hp) x = 2 * rnd() - 1;
  double *th, *p, *out;
  cudaMalloc(&th, sizeof(double) * nsys * N); cudaMalloc(&p, sizeof(double) * nsys * N); cudaMalloc(&out, sizeof(double) * nsys * N);
  cudaMemcpy(th, hth.data(), sizeof(double) * nsys * N, cudaMemcpyHostToDevice);
  cudaMemcpy(p, hp.data(), sizeof(double) * nsys * N, cudaMemcpyHostToDevice);
  printf("12 x 12 chain mass matrix: assemble + LDL^T + solve, %d systems x %d dependent solves\n", nsys, reps);
  run("thread per system", solve_thread, 1, th, p, out, nsys, reps, hth, hp);
  run("4 lanes per system (shfl)", solve_lanes<4>, 4, th, p, out, nsys, reps, hth, hp);
  run("8 lanes per system (shfl)", solve_lanes<8>, 8, th, p, out, nsys, reps, hth, hp);
  run("16 lanes per system (shfl)", solve_lanes<16>, 16, th, p, out, nsys, reps, hth, hp);
  return 0;
}
