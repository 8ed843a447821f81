// ensemble_main.cpp — drives the multi-GPU ensemble of the C ABI (hb_ensemble_*) from a plain C++ program: no Python, no
// torchrun, one process for all GPUs — the way a compiled (Haskell / C++) host uses it.  Used by tests/test_gpu_parity.py
// (every GPU the box has) and by profiles/gpu_round_*.sh for the BASELINE configs[3] measurement:
//   ensemble_main <ndev> <system id> <N> <steps> [check]
// Triple pendulum (id 6), N initial conditions split over ndev GPUs, `steps` one-step RK4 launches, ONE all-gather.
// `check`: every gathered Phase is compared with a single-GPU hb_batch_step recomputation (bit for bit) and the gathered
// copies of all devices are compared with each other.  Prints one JSON line.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../include/hamilton_b200.h"

#define HB(call) do { hb_status s_ = (call); if (s_) { std::fprintf(stderr, "%s -> %d: %s\n", #call, (int)s_, hb_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  const int ndev = argc > 1 ? std::atoi(argv[1]) : 1;
  const int sid = argc > 2 ? std::atoi(argv[2]) : HB_SYS_TRIPLE_PENDULUM;
  const long long N = argc > 3 ? std::atoll(argv[3]) : 1 << 16;
  const int steps = argc > 4 ? std::atoi(argv[4]) : 10;
  const bool check = argc > 5 && std::strcmp(argv[5], "check") == 0;
  hb_system* sys = nullptr;
  HB(hb_system_builtin((hb_builtin)sid, nullptr, 0, &sys));
  int32_t m = 0, n = 0;
  HB(hb_system_dims(sys, &m, &n));
  const int D = 2 * n;
  std::vector<double> lo(D), hi(D);
  for (int c = 0; c < D; c++) { lo[c] = c < n ? -M_PI : -1.0; hi[c] = c < n ? M_PI : 1.0; }
  hb_ensemble* ens = nullptr;
  HB(hb_ensemble_create(sys, ndev, nullptr, N, &ens));
  HB(hb_ensemble_init_random(ens, 0x48414D49ULL, lo.data(), hi.data()));
  double warm_ms = 0, step_ms = 0, gather_ms = 0, gather2_ms = 0;
  HB(hb_ensemble_step(ens, HB_INTEG_RK4, 0.01, 1, 3, &warm_ms));        // warm-up launches (3 more steps of the same trajectories)
  HB(hb_ensemble_gather(ens, nullptr, &gather_ms));                     // first gather: allocates + registers the buffers
  const auto t0 = std::chrono::steady_clock::now();
  HB(hb_ensemble_step(ens, HB_INTEG_RK4, 0.01, 1, steps, &step_ms));
  std::vector<double> all(check ? (size_t)N * D : 0);
  HB(hb_ensemble_gather(ens, check ? all.data() : nullptr, &gather2_ms));
  const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  long long bad = 0, flagged = 0;
  if (check) {
    // recompute everything on device 0 alone through the batch entry point and compare bit for bit
    HB(hb_set_device(0));
    double *y = nullptr, *y2 = nullptr;
    cudaMalloc((void**)&y, sizeof(double) * N * D);
    cudaMalloc((void**)&y2, sizeof(double) * N * D);
    HB(hb_batch_init_random(sys, 0x48414D49ULL, 0, N, HB_LAYOUT_AOS, lo.data(), hi.data(), y, nullptr));
    for (int s = 0; s < steps + 3; s++) { HB(hb_batch_step(sys, HB_INTEG_RK4, 0.01, 1, N, HB_LAYOUT_AOS, HB_MEM_DEVICE, y, y2, nullptr, nullptr)); std::swap(y, y2); }
    std::vector<double> ref((size_t)N * D);
    cudaMemcpy(ref.data(), y, sizeof(double) * N * D, cudaMemcpyDeviceToHost);
    for (size_t k = 0; k < ref.size(); k++) bad += std::memcmp(&ref[k], &all[k], 8) != 0;
    // every device's gathered copy equals device 0's
    std::vector<double> other((size_t)N * D);
    for (int g = 1; g < ndev; g++) {
      double* p = nullptr;
      HB(hb_ensemble_gathered(ens, g, &p));
      cudaMemcpy(other.data(), p, sizeof(double) * N * D, cudaMemcpyDefault);
      bad += std::memcmp(other.data(), all.data(), sizeof(double) * N * D) != 0;
    }
    std::vector<int32_t> fl(N);
    HB(hb_ensemble_flags(ens, fl.data()));
    for (int32_t f : fl) flagged += f != 0;
    cudaFree(y); cudaFree(y2);
  }
  const double total = (double)N * steps;
  std::printf("{\"ndev\": %d, \"system\": %d, \"N\": %lld, \"steps\": %d, \"step_ms\": %.4f, \"gather_first_ms\": %.4f, \"gather_ms\": %.4f, "
              "\"wall_ms\": %.3f, \"steps_per_s\": %.6e, \"steps_per_s_with_gather\": %.6e, \"gather_bus_GBps\": %.2f, \"mismatches\": %lld, \"flagged\": %lld}\n",
              ndev, sid, N, steps, step_ms, gather_ms, gather2_ms, wall_ms, total / (step_ms * 1e-3), total / ((step_ms + gather2_ms) * 1e-3),
              ndev > 1 ? (double)N * D * 8 * (ndev - 1) / ndev / (gather2_ms * 1e-3) / 1e9 : 0.0, bad, flagged);
  hb_ensemble_free(ens);
  hb_system_free(sys);
  return bad ? 2 : 0;
}
