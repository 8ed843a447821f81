// ensemble.cpp — multi-GPU ensembles behind the C ABI (include/hamilton_b200.h, hb_ensemble_*).
//
// N independent trajectories of one System, block-split over the GPUs of this process: device g owns
// [g*N/G, (g+1)*N/G) (independence: src/Numeric/Hamilton.hs:390-399).  Every call runs one host
// thread per device; stepping is hb_batch_step on each shard with no communication; the collection is ONE
// ncclAllGather over NVLink on registered buffers (ragged shards: one grouped ncclBroadcast per shard).
// NCCL is loaded with dlopen at the first ensemble creation (no link-time dependency; with ndev == 1 it is not needed).
// Built on the public batch entry points of this library — no second copy of the launch logic.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hamilton_b200.h"

extern "C" hb_status hb_internal_fail(hb_status s, const char* msg);   // runtime.cpp: sets the caller's thread-local error

namespace {

// ---- the few NCCL entry points used, bound at run time (types as in nccl.h 2.x) ---------------------------------
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;          // ncclSuccess = 0
enum { kNcclFloat64 = 8 };         // ncclDataType_t ncclFloat64
struct Nccl {
  void* h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*CommRegister)(ncclComm_t, void*, size_t, void**) = nullptr;      // optional (NCCL >= 2.19)
  ncclResult_t (*CommDeregister)(ncclComm_t, void*) = nullptr;
  std::string err;
  bool load() {
    if (h) return true;
    for (const char* n : {"libnccl.so.2", "libnccl.so"}) { h = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (h) break; }
    if (!h) { err = "cannot dlopen libnccl.so.2 (needed for ensembles over more than one GPU)"; return false; }
#define SYM(x, req) x = (decltype(x))dlsym(h, "nccl" #x); if (req && !x) { err = "libnccl lacks nccl" #x; h = nullptr; return false; }
    SYM(CommInitAll, 1) SYM(CommDestroy, 1) SYM(AllGather, 1) SYM(Broadcast, 1) SYM(GroupStart, 1) SYM(GroupEnd, 1) SYM(GetErrorString, 1)
    SYM(GetVersion, 0) SYM(CommRegister, 0) SYM(CommDeregister, 0)
#undef SYM
    return true;
  }
};
Nccl g_nccl;
std::mutex g_nccl_mu;

struct Shard {
  int device = 0;
  int64_t first = 0, n = 0;
  cudaStream_t st = nullptr;
  double* buf[2] = {nullptr, nullptr};
  int cur = 0;
  int32_t* flags = nullptr;
  double* gathered = nullptr;
  ncclComm_t comm = nullptr;
  void* reg[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

}  // namespace

struct hb_ensemble {
  const hb_system* sys = nullptr;
  int n = 0, m = 0, D = 0, ndev = 0;
  int64_t N = 0;
  bool equal = true;          // all shards the same size: ncclAllGather; else grouped ncclBroadcast
  std::vector<Shard> sh;
};

namespace {

// Runs f(g) on one host thread per device (cudaSetDevice done), returns the first failure.
template <class F>
hb_status on_devices(hb_ensemble* e, F f) {
  std::vector<hb_status> rc(e->ndev, HB_OK);
  std::vector<std::string> msg(e->ndev);
  auto body = [&](int g) {
    cudaError_t ce = cudaSetDevice(e->sh[g].device);
    if (ce != cudaSuccess) { rc[g] = HB_ERR_CUDA; msg[g] = std::string("cudaSetDevice: ") + cudaGetErrorString(ce); return; }
    rc[g] = f(g, msg[g]);
    if (rc[g] != HB_OK && msg[g].empty()) msg[g] = hb_last_error();
  };
  if (e->ndev == 1) body(0);
  else {
    std::vector<std::thread> th;
    for (int g = 0; g < e->ndev; g++) th.emplace_back(body, g);
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < e->ndev; g++)
    if (rc[g] != HB_OK) return hb_internal_fail(rc[g], ("device " + std::to_string(e->sh[g].device) + ": " + msg[g]).c_str());
  return HB_OK;
}
#define CUE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { m = std::string(#call ": ") + cudaGetErrorString(e_); return (hb_status)HB_ERR_CUDA; } } while (0)
#define NCE(call) do { ncclResult_t r_ = (call); if (r_ != 0) { m = std::string(#call ": ") + g_nccl.GetErrorString(r_); return (hb_status)HB_ERR_CUDA; } } while (0)

}  // namespace

extern "C" {

hb_status hb_ensemble_create(const hb_system* sys, int32_t ndev, const int32_t* devices, int64_t N, hb_ensemble** out) {
  if (!out) return hb_internal_fail(HB_ERR_INVALID, "null out");
  *out = nullptr;
  if (!sys || ndev < 1 || N < 0) return hb_internal_fail(HB_ERR_INVALID, "bad ensemble arguments");
  int32_t have = 0;
  hb_status rc = hb_device_count(&have);
  if (rc) return rc;
  for (int g = 0; g < ndev; g++) {
    const int d = devices ? devices[g] : g;
    if (d < 0 || d >= have) return hb_internal_fail(HB_ERR_INVALID, "ensemble device ordinal out of range");
    for (int k = 0; k < g; k++) if ((devices ? devices[k] : k) == d) return hb_internal_fail(HB_ERR_INVALID, "ensemble devices must be distinct");
  }
  hb_ensemble* e = new hb_ensemble();
  e->sys = sys; e->ndev = ndev; e->N = N;
  int32_t sm = 0, sn = 0;
  hb_system_dims(sys, &sm, &sn);
  e->m = sm; e->n = sn; e->D = 2 * sn;
  e->sh.resize(ndev);
  for (int g = 0; g < ndev; g++) {
    Shard& s = e->sh[g];
    s.device = devices ? devices[g] : g;
    s.first = N * g / ndev;
    s.n = N * (g + 1) / ndev - s.first;
    if (s.n != e->sh[0].n) e->equal = false;
  }
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  rc = on_devices(e, [&](int g, std::string& m) -> hb_status {
    Shard& s = e->sh[g];
    CUE(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    CUE(cudaEventCreate(&s.e0));
    CUE(cudaEventCreate(&s.e1));
    const size_t bytes = (size_t)(s.n > 0 ? s.n : 1) * e->D * sizeof(double);
    CUE(cudaMalloc((void**)&s.buf[0], bytes));
    CUE(cudaMalloc((void**)&s.buf[1], bytes));
    CUE(cudaMalloc((void**)&s.flags, sizeof(int32_t) * (size_t)(s.n > 0 ? s.n : 1)));
    CUE(cudaMemsetAsync(s.buf[0], 0, bytes, s.st));
    CUE(cudaMemsetAsync(s.flags, 0, sizeof(int32_t) * (size_t)(s.n > 0 ? s.n : 1), s.st));
    CUE(cudaStreamSynchronize(s.st));
    return HB_OK;
  });
  if (!rc && ndev > 1) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!g_nccl.load()) rc = hb_internal_fail(HB_ERR_UNSUPPORTED, g_nccl.err.c_str());
    else {
      std::vector<ncclComm_t> comms(ndev);
      std::vector<int> devs(ndev);
      for (int g = 0; g < ndev; g++) devs[g] = e->sh[g].device;
      ncclResult_t r = g_nccl.CommInitAll(comms.data(), ndev, devs.data());
      if (r != 0) rc = hb_internal_fail(HB_ERR_CUDA, (std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r)).c_str());
      else for (int g = 0; g < ndev; g++) e->sh[g].comm = comms[g];
    }
  }
  cudaSetDevice(cur_dev);
  if (rc) { hb_ensemble_free(e); return rc; }
  *out = e;
  return HB_OK;
}

void hb_ensemble_free(hb_ensemble* e) {
  if (!e) return;
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  for (auto& s : e->sh) {
    cudaSetDevice(s.device);
    if (s.st) cudaStreamSynchronize(s.st);
    if (s.comm) {
      if (g_nccl.CommDeregister) for (void* r : s.reg) if (r) g_nccl.CommDeregister(s.comm, r);
      g_nccl.CommDestroy(s.comm);
    }
    for (double* b : s.buf) if (b) cudaFree(b);
    if (s.flags) cudaFree(s.flags);
    if (s.gathered) cudaFree(s.gathered);
    if (s.e0) cudaEventDestroy(s.e0);
    if (s.e1) cudaEventDestroy(s.e1);
    if (s.st) cudaStreamDestroy(s.st);
  }
  cudaSetDevice(cur_dev);
  delete e;
}

hb_status hb_ensemble_dims(const hb_ensemble* e, int32_t* ndev, int64_t* N, int64_t* first) {
  if (!e) return hb_internal_fail(HB_ERR_INVALID, "null ensemble");
  if (ndev) *ndev = e->ndev;
  if (N) *N = e->N;
  if (first) { for (int g = 0; g < e->ndev; g++) first[g] = e->sh[g].first; first[e->ndev] = e->N; }
  return HB_OK;
}

hb_status hb_ensemble_init_random(hb_ensemble* e, uint64_t seed, const double* lo, const double* hi) {
  if (!e || !lo || !hi) return hb_internal_fail(HB_ERR_INVALID, "null argument");
  return on_devices(e, [&](int g, std::string& m) -> hb_status {
    Shard& s = e->sh[g];
    hb_status rc = hb_batch_init_random(e->sys, seed, s.first, s.n, HB_LAYOUT_AOS, lo, hi, s.buf[s.cur], s.st);
    if (rc) return rc;
    CUE(cudaStreamSynchronize(s.st));
    return HB_OK;
  });
}

hb_status hb_ensemble_upload(hb_ensemble* e, const double* y_host) {
  if (!e || !y_host) return hb_internal_fail(HB_ERR_INVALID, "null argument");
  return on_devices(e, [&](int g, std::string& m) -> hb_status {
    Shard& s = e->sh[g];
    CUE(cudaMemcpyAsync(s.buf[s.cur], y_host + (size_t)s.first * e->D, (size_t)s.n * e->D * sizeof(double), cudaMemcpyHostToDevice, s.st));
    CUE(cudaStreamSynchronize(s.st));
    return HB_OK;
  });
}

hb_status hb_ensemble_step(hb_ensemble* e, hb_integrator integ, double dt, int32_t nsteps, int32_t launches, double* gpu_ms) {
  if (!e) return hb_internal_fail(HB_ERR_INVALID, "null ensemble");
  if (launches < 0) return hb_internal_fail(HB_ERR_INVALID, "negative launches");
  std::vector<float> ms(e->ndev, 0.f);
  hb_status rc = on_devices(e, [&](int g, std::string& m) -> hb_status {
    Shard& s = e->sh[g];
    CUE(cudaEventRecord(s.e0, s.st));
    for (int l = 0; l < launches; l++) {
      hb_status r = hb_batch_step(e->sys, integ, dt, nsteps, s.n, HB_LAYOUT_AOS, HB_MEM_DEVICE, s.buf[s.cur], s.buf[s.cur ^ 1], s.flags, s.st);
      if (r) return r;
      s.cur ^= 1;
    }
    CUE(cudaEventRecord(s.e1, s.st));
    CUE(cudaStreamSynchronize(s.st));
    CUE(cudaEventElapsedTime(&ms[g], s.e0, s.e1));
    return HB_OK;
  });
  if (gpu_ms) { float mx = 0.f; for (float x : ms) mx = x > mx ? x : mx; *gpu_ms = mx; }
  return rc;
}

hb_status hb_ensemble_gather(hb_ensemble* e, double* y_host, double* gpu_ms) {
  if (!e) return hb_internal_fail(HB_ERR_INVALID, "null ensemble");
  std::vector<float> ms(e->ndev, 0.f);
  const size_t all_bytes = (size_t)(e->N > 0 ? e->N : 1) * e->D * sizeof(double);
  hb_status rc = on_devices(e, [&](int g, std::string& m) -> hb_status {
    Shard& s = e->sh[g];
    if (!s.gathered) {
      CUE(cudaMalloc((void**)&s.gathered, all_bytes));
      if (s.comm && g_nccl.CommRegister) {   // registered user buffers: zero-copy NVLink/NVLS paths instead of staging
        const size_t sb = (size_t)(s.n > 0 ? s.n : 1) * e->D * sizeof(double);
        g_nccl.CommRegister(s.comm, s.buf[0], sb, &s.reg[0]);
        g_nccl.CommRegister(s.comm, s.buf[1], sb, &s.reg[1]);
        g_nccl.CommRegister(s.comm, s.gathered, all_bytes, &s.reg[2]);
      }
    }
    CUE(cudaEventRecord(s.e0, s.st));
    if (e->ndev == 1) {
      CUE(cudaMemcpyAsync(s.gathered, s.buf[s.cur], (size_t)s.n * e->D * sizeof(double), cudaMemcpyDeviceToDevice, s.st));
    } else if (e->equal) {
      NCE(g_nccl.AllGather(s.buf[s.cur], s.gathered, (size_t)s.n * e->D, kNcclFloat64, s.comm, s.st));
    } else {   // ragged shards: one broadcast per shard, fused into one group
      NCE(g_nccl.GroupStart());
      for (int r = 0; r < e->ndev; r++) {
        const Shard& root = e->sh[r];
        ncclResult_t nr = g_nccl.Broadcast(s.buf[s.cur], s.gathered + (size_t)root.first * e->D, (size_t)root.n * e->D, kNcclFloat64, r, s.comm, s.st);
        if (nr != 0) { g_nccl.GroupEnd(); m = std::string("ncclBroadcast: ") + g_nccl.GetErrorString(nr); return (hb_status)HB_ERR_CUDA; }
      }
      NCE(g_nccl.GroupEnd());
    }
    CUE(cudaEventRecord(s.e1, s.st));
    if (g == 0 && y_host) CUE(cudaMemcpyAsync(y_host, s.gathered, (size_t)e->N * e->D * sizeof(double), cudaMemcpyDeviceToHost, s.st));
    CUE(cudaStreamSynchronize(s.st));
    CUE(cudaEventElapsedTime(&ms[g], s.e0, s.e1));
    return HB_OK;
  });
  if (gpu_ms) { float mx = 0.f; for (float x : ms) mx = x > mx ? x : mx; *gpu_ms = mx; }
  return rc;
}

hb_status hb_ensemble_shard(const hb_ensemble* e, int32_t g, double** y_device, int64_t* n_shard) {
  if (!e || g < 0 || g >= e->ndev) return hb_internal_fail(HB_ERR_INVALID, "bad ensemble / device index");
  if (y_device) *y_device = e->sh[g].buf[e->sh[g].cur];
  if (n_shard) *n_shard = e->sh[g].n;
  return HB_OK;
}
hb_status hb_ensemble_gathered(const hb_ensemble* e, int32_t g, double** y_device) {
  if (!e || g < 0 || g >= e->ndev || !y_device) return hb_internal_fail(HB_ERR_INVALID, "bad ensemble / device index");
  if (!e->sh[g].gathered) return hb_internal_fail(HB_ERR_INVALID, "hb_ensemble_gather has not run yet");
  *y_device = e->sh[g].gathered;
  return HB_OK;
}
hb_status hb_ensemble_flags(hb_ensemble* e, int32_t* flags_host) {
  if (!e || !flags_host) return hb_internal_fail(HB_ERR_INVALID, "null argument");
  return on_devices(e, [&](int g, std::string& m) -> hb_status {
    Shard& s = e->sh[g];
    CUE(cudaMemcpyAsync(flags_host + s.first, s.flags, sizeof(int32_t) * (size_t)s.n, cudaMemcpyDeviceToHost, s.st));
    CUE(cudaStreamSynchronize(s.st));
    return HB_OK;
  });
}

}  // extern "C"
