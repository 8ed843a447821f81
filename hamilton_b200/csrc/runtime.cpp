// runtime.cpp — C ABI implementation (include/hamilton_b200.h): system objects, NVRTC JIT of tape
// systems, kernel dispatch, host<->device staging.  Host-only C++; the device code lives in
// engine/hb_engine.cuh (hand-written) + the per-system derivative structs printed by sysgen.cpp.
//
// There is deliberately no CPU implementation of any compute entry point in this library.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/hamilton_b200.h"
#include "builtins.hpp"
#include "sysgen.hpp"

// ---- mirrors of the device-side declarations in engine/hb_engine.cuh (kept in sync by a static_assert in aot_kernels.cu)
#define HB_MAXP 64
struct HbKArgs {
  const double* in; double* out; int* flags; const double* ts;
  long long N; double dt; double dt6; double dth; int nsteps; int layout; int s; int substeps;
  unsigned long long seed; long long first;
  int host_io; int contiguous;
  double prm[HB_MAXP];
};
enum { K_STEP_RK4 = 0, K_STEP_RKF45, K_EVOLVE_RK4, K_EVOLVE_RKF45, K_HAM_EQS, K_TO_PHASE, K_FROM_PHASE, K_ENERGIES, K_UPOS, K_COUNT };
static const char* const KERNEL_KINDS[K_COUNT] = {"step_rk4", "step_rkf45", "evolve_rk4", "evolve_rkf45", "ham_eqs",
                                                  "to_phase", "from_phase", "energies", "upos"};
#define HB_BLOCK 128
#define HB_BLOCK_OF(NCOORD) 128   // large systems (n >= HB_BIG_N); must match engine/hb_engine.cuh
#define HB_BIG_N 8
#define HB_WSTORE_MAXD 8   // must match engine/hb_engine.cuh
#define HB_DYN_DOUBLES(NCOORD, NE_) ((NCOORD) >= HB_BIG_N ? 3 * 2 * (NCOORD) + (NE_) : 0)
#define HB_MAXBLOCK_OF(NCOORD) HB_BLOCK_OF(NCOORD)
#define HB_TAB_BYTES 33408   // must match engine/hb_engine.cuh
#define HB_MAX_LAUNCH_N ((int64_t)0x7C000000)   // 2^31 - 2^26: i + (trajectories per round) stays below 2^32 in the kernels

// from aot_kernels.cu
extern "C" const void* hb_aot_kernel(int builtin, int kernel_id);
extern "C" const void* hb_aot_init_random(void);
extern "C" int hb_aot_dyn_doubles(int builtin);
extern "C" size_t hb_aot_kargs_size(void);
// from gen/engine_embed.inc (the engine header as a string, for NVRTC)
extern const char hb_engine_src[];
extern const char hb_sincos_tab_src[];

namespace {

thread_local std::string g_err;
hb_status fail(hb_status s, const std::string& msg) { g_err = msg; return s; }
hb_status cuda_fail(cudaError_t e, const char* what) {
  // errors that mean "this machine cannot run CUDA at all"
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError ||
      e == cudaErrorNotSupported || e == cudaErrorSystemDriverMismatch || e == cudaErrorSystemNotReady)
    return fail(HB_ERR_NO_DEVICE, std::string(what) + ": " + cudaGetErrorString(e) + " (no CUDA device; this library has no CPU fallback)");
  return fail(HB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

hb_status need_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(e, "cudaGetDeviceCount"); }
  if (n <= 0) return fail(HB_ERR_NO_DEVICE, "no CUDA device; this library has no CPU fallback");
  return HB_OK;
}

// ------------------------------------------------------------------------------ NVRTC -------
struct Nvrtc {
  void* h = nullptr;
  decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
  decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
  decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
  decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
  decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
  decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
  decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
  decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
  decltype(&nvrtcVersion) Version = nullptr;
  std::string err;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.13"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (h) break; }
    if (!h) { err = "cannot dlopen libnvrtc (needed to compile tape systems)"; return false; }
#define SYM(x) x = (decltype(x))dlsym(h, "nvrtc" #x); if (!x) { err = "libnvrtc lacks nvrtc" #x; return false; }
    SYM(CreateProgram) SYM(CompileProgram) SYM(GetCUBINSize) SYM(GetCUBIN) SYM(GetProgramLogSize) SYM(GetProgramLog)
    SYM(DestroyProgram) SYM(GetErrorString) SYM(Version)
#undef SYM
    return true;
  }
};
Nvrtc g_nvrtc;
std::mutex g_nvrtc_mu;

std::string jit_arch() {
  if (const char* e = std::getenv("HB_JIT_ARCH")) return e;
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
    char buf[32];
    const bool arch_specific = p.major >= 9;   // sm_90a / sm_100a / sm_103a
    std::snprintf(buf, sizeof buf, "sm_%d%d%s", p.major, p.minor, arch_specific ? "a" : "");
    return buf;
  }
  cudaGetLastError();
  return "sm_100a";   // the target this library is written for
}

// ---- on-disk cache of NVRTC output ------------------------------------------------------------
// mkSystem is a compiler here (1-2 s for the reference's examples, over a minute for a 12-coordinate chain), and a host
// program builds the same Systems on every start.  The cubin is therefore kept under $HB_JIT_CACHE_DIR (default
// $XDG_CACHE_HOME/hamilton_b200 or ~/.cache/hamilton_b200), keyed by a 128-bit hash of everything that determines it:
// the engine header, the generated translation unit, the architecture, the options and the NVRTC version.
// HB_JIT_CACHE=0 disables it.  Files are written to a temporary name and renamed, so concurrent processes are safe.
struct Hash128 {
  unsigned long long a = 0xcbf29ce484222325ULL, b = 0x84222325cbf29ce4ULL;
  void add(const void* p, size_t n) {
    const unsigned char* c = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) {
      a = (a ^ c[i]) * 0x100000001b3ULL;                       // FNV-1a
      b = (b + c[i] + 0x9e3779b97f4a7c15ULL) * 0xff51afd7ed558ccdULL; b ^= b >> 32;   // an unrelated second mix
    }
    const unsigned long long len = n;
    for (int i = 0; i < 8; i++) { a = (a ^ ((len >> (8 * i)) & 0xff)) * 0x100000001b3ULL; }
  }
  void add(const std::string& s) { add(s.data(), s.size()); }
  std::string hex() const { char buf[40]; std::snprintf(buf, sizeof buf, "%016llx%016llx", a, b); return buf; }
};

std::string jit_cache_dir() {
  const char* off = std::getenv("HB_JIT_CACHE");
  if (off && off[0] == '0') return "";
  std::string d;
  if (const char* e = std::getenv("HB_JIT_CACHE_DIR")) d = e;
  else if (const char* x = std::getenv("XDG_CACHE_HOME")) d = std::string(x) + "/hamilton_b200";
  else if (const char* h = std::getenv("HOME")) d = std::string(h) + "/.cache/hamilton_b200";
  if (d.empty()) return "";
  for (size_t i = 1; i <= d.size(); i++)                       // mkdir -p
    if (i == d.size() || d[i] == '/') { std::string sub = d.substr(0, i); mkdir(sub.c_str(), 0777); }
  struct stat st;
  if (stat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode) || access(d.c_str(), W_OK) != 0) return "";
  return d;
}

bool read_file(const std::string& path, std::vector<char>& out) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  bool ok = n > 4;
  if (ok) { out.resize((size_t)n); ok = std::fread(out.data(), 1, (size_t)n, f) == (size_t)n; }
  std::fclose(f);
  return ok && out[0] == 0x7f && out[1] == 'E' && out[2] == 'L' && out[3] == 'F';   // a cubin is an ELF image
}

void write_file_atomic(const std::string& path, const std::vector<char>& data) {
  char suffix[64];
  std::snprintf(suffix, sizeof suffix, ".tmp.%ld.%p", (long)getpid(), (const void*)&data);
  const std::string tmp = path + suffix;
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) return;
  const bool ok = std::fwrite(data.data(), 1, data.size(), f) == data.size();
  if (std::fclose(f) != 0 || !ok || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}

bool nvrtc_compile_uncached(const std::string& src, const std::string& arch, std::vector<char>& cubin, std::string& log);

bool nvrtc_compile(const std::string& src, const std::string& arch, std::vector<char>& cubin, std::string& log) {
  const std::string dir = jit_cache_dir();
  std::string path;
  if (!dir.empty()) {
    Hash128 h;
    h.add(hb_engine_src, std::strlen(hb_engine_src));
    h.add(hb_sincos_tab_src, std::strlen(hb_sincos_tab_src));
    h.add(src);
    h.add(arch);
    if (const char* e = std::getenv("HB_JIT_DEFINES")) h.add(e, std::strlen(e));
    int vmaj = 0, vmin = 0;
    { std::lock_guard<std::mutex> lk(g_nvrtc_mu); if (g_nvrtc.load() && g_nvrtc.Version) g_nvrtc.Version(&vmaj, &vmin); }
    h.add(&vmaj, sizeof vmaj); h.add(&vmin, sizeof vmin);
    path = dir + "/" + h.hex() + ".cubin";
    if (read_file(path, cubin)) return true;
    cubin.clear();
  }
  if (!nvrtc_compile_uncached(src, arch, cubin, log)) return false;
  if (!path.empty()) write_file_atomic(path, cubin);
  return true;
}

bool nvrtc_compile_uncached(const std::string& src, const std::string& arch, std::vector<char>& cubin, std::string& log) {
  std::lock_guard<std::mutex> lk(g_nvrtc_mu);
  if (!g_nvrtc.load()) { log = g_nvrtc.err; return false; }
  nvrtcProgram prog;
  const char* hdr_src[] = {hb_engine_src, hb_sincos_tab_src};
  const char* hdr_name[] = {"hb_engine.cuh", "hb_sincos_tab.cuh"};
  nvrtcResult r = g_nvrtc.CreateProgram(&prog, src.c_str(), "hb_jit_system.cu", 2, hdr_src, hdr_name);
  if (r != NVRTC_SUCCESS) { log = std::string("nvrtcCreateProgram: ") + g_nvrtc.GetErrorString(r); return false; }
  std::string a = "--gpu-architecture=" + arch;
  std::vector<std::string> extra;   // HB_JIT_DEFINES="HB_MINB_RK4=8,FOO=1": tuning experiments without a rebuild
  if (const char* e = std::getenv("HB_JIT_DEFINES")) {
    std::string s(e); size_t p0 = 0;
    while (p0 <= s.size()) { size_t p1 = s.find(',', p0); if (p1 == std::string::npos) p1 = s.size(); if (p1 > p0) extra.push_back("-D" + s.substr(p0, p1 - p0)); p0 = p1 + 1; }
  }
  std::vector<const char*> opts = {a.c_str(), "--std=c++17", "-lineinfo", "--fmad=true", "-default-device"};
  for (auto& x : extra) opts.push_back(x.c_str());
  r = g_nvrtc.CompileProgram(prog, (int)opts.size(), opts.data());
  size_t ls = 0;
  g_nvrtc.GetProgramLogSize(prog, &ls);
  if (ls > 1) { log.resize(ls); g_nvrtc.GetProgramLog(prog, &log[0]); }
  if (r != NVRTC_SUCCESS) { log = std::string("nvrtcCompileProgram: ") + g_nvrtc.GetErrorString(r) + "\n" + log; g_nvrtc.DestroyProgram(&prog); return false; }
  size_t cs = 0;
  g_nvrtc.GetCUBINSize(prog, &cs);
  cubin.resize(cs);
  g_nvrtc.GetCUBIN(prog, cubin.data());
  g_nvrtc.DestroyProgram(&prog);
  return true;
}

}  // namespace

// --------------------------------------------------------------------------- hb_system -------
struct hb_system {
  int m = 0, n = 0;
  int builtin = -1;                    // hb_builtin id or -1 (tape / JIT)
  bool baked = false;                  // built-in with the default parameters: use the literal-specialised kernels
  std::vector<double> params;          // tape-level runtime parameters (<= HB_MAXP)
  int dyn_doubles = 0;                 // dynamic shared memory per thread (doubles) every kernel of this system is launched with
  int rhs_cost = 0;                    // system compiler's cost model of one hamEqs evaluation
  bool heavy = false;                  // Sys::HEAVY: one RK4 step is issue-bound, not HBM-bound (launch-shape heuristic)
  bool trig = false;                   // Sys::TRIG (sin / cos / exp on the tapes): the kernels stage the table image in dynamic shared memory
  double intensity = 0;                // issue clocks / HBM clocks of one RK4 step (system compiler's estimate)
  std::string source;                  // generated Sys struct
  // JIT: small systems compile all kernels in one NVRTC program at creation (cubins[K_COUNT] shared slot 0);
  // large ones compile each kernel on first use (a 12-coordinate chain takes ~10 s per kernel).
  hb::GeneratedSystem gen;
  std::string arch;
  bool lazy = false, no_code = false;
  std::vector<char> cubins[K_COUNT];
  cudaLibrary_t libs[K_COUNT] = {nullptr};
  const void* jit_kernels[K_COUNT] = {nullptr};
  std::mutex mu;

  hb_status kernel(int kid, const void** fn) {
    if (builtin >= 0) {
      *fn = hb_aot_kernel(builtin + (baked ? HB_SYS__COUNT : 0), kid);
      return *fn ? HB_OK : fail(HB_ERR_INVALID, "no such AOT kernel");
    }
    std::lock_guard<std::mutex> lk(mu);
    if (no_code) return fail(HB_ERR_COMPILE, "system was created with HB_JIT_SKIP_COMPILE: no device code");
    if (!jit_kernels[kid]) {
      const int slot = lazy ? kid : 0;
      if (cubins[slot].empty()) {   // lazy: compile just this kernel now
        std::string log;
        if (!nvrtc_compile(hb::jit_translation_unit(gen, "hbk", KERNEL_KINDS[kid]), arch, cubins[slot], log)) return fail(HB_ERR_COMPILE, log);
      }
      if (!libs[slot] && cudaLibraryLoadData(&libs[slot], cubins[slot].data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
        // the image may have come from the disk cache (another driver, a damaged file): compile afresh once and retry
        cudaGetLastError();
        libs[slot] = nullptr;
        std::string log;
        const std::string tu = lazy ? hb::jit_translation_unit(gen, "hbk", KERNEL_KINDS[kid]) : hb::jit_translation_unit(gen, "hbk");
        if (!nvrtc_compile_uncached(tu, arch, cubins[slot], log)) return fail(HB_ERR_COMPILE, log);
        CU(cudaLibraryLoadData(&libs[slot], cubins[slot].data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
      }
      cudaKernel_t kh;
      std::string nm = std::string("hbk_") + KERNEL_KINDS[kid];
      CU(cudaLibraryGetKernel(&kh, libs[slot], nm.c_str()));
      jit_kernels[kid] = (const void*)kh;
    }
    *fn = jit_kernels[kid];
    return HB_OK;
  }
};

namespace {

// Identity of one captured host-staging pipeline (run_batch): kernel, caller buffers, device scratch and every argument.
struct HostPipeKey {
  const void *fn, *in, *out, *flags, *din, *dout, *dfl;
  int64_t N, chunks;
  int in_d, out_d, dyn, pad_;
  HbKArgs a;
};

// Device scratch for HB_MEM_HOST calls: per-thread, grow-only, with a private stream.
struct Scratch {
  // small most-recently-used cache of instantiated chunk-pipeline graphs
  struct Pipe { HostPipeKey key; cudaGraphExec_t exec; };
  std::vector<Pipe> pipes;
  cudaGraphExec_t find_pipe(const HostPipeKey& k) {
    for (size_t i = 0; i < pipes.size(); i++)
      if (std::memcmp(&pipes[i].key, &k, sizeof k) == 0) {
        Pipe hit = pipes[i];
        pipes.erase(pipes.begin() + i);
        pipes.insert(pipes.begin(), hit);
        return hit.exec;
      }
    return nullptr;
  }
  void add_pipe(const HostPipeKey& k, cudaGraphExec_t e) {
    if (pipes.size() >= 8) { cudaGraphExecDestroy(pipes.back().exec); pipes.pop_back(); }
    pipes.insert(pipes.begin(), Pipe{k, e});
  }
  void drop_pipes() {
    for (auto& p : pipes) cudaGraphExecDestroy(p.exec);
    pipes.clear();
  }
  // thread exit: a -threaded Haskell host makes HOST calls from many short-lived OS threads; give everything back
  ~Scratch() {
    if (device < 0) return;
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return; }   // (runtime already shut down at process exit)
    cudaSetDevice(device);
    drop_pipes();
    for (auto e : events) cudaEventDestroy(e);
    for (auto& st : streams) if (st) cudaStreamDestroy(st);
    for (int i = 0; i < 3; i++) if (p[i]) cudaFree(p[i]);
    cudaSetDevice(cur);
    cudaGetLastError();
  }
  void* p[3] = {nullptr, nullptr, nullptr};
  size_t cap[3] = {0, 0, 0};
  cudaStream_t stream = nullptr;      // == streams[0]
  cudaStream_t streams[3] = {nullptr, nullptr, nullptr};   // chunk pipeline: H2D / kernel / D2H of neighbouring chunks overlap
  int device = -1;
  std::vector<cudaEvent_t> events;    // chunk pipeline: upload-done / kernel-done per chunk
  hb_status need_events(int n) {
    while ((int)events.size() < n) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      events.push_back(e);
    }
    return HB_OK;
  }
  hb_status get(int slot, size_t bytes, void** out) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev != device) {   // device switched on this thread: drop the old buffers
      for (int i = 0; i < 3; i++) { if (p[i]) { cudaSetDevice(device); cudaFree(p[i]); cudaSetDevice(dev); } p[i] = nullptr; cap[i] = 0; }
      for (auto& st : streams) if (st) { cudaStreamDestroy(st); st = nullptr; }
      for (auto e : events) cudaEventDestroy(e);
      events.clear();
      drop_pipes();
      stream = nullptr;
      device = dev;
    }
    if (!stream) {
      for (auto& st : streams) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      stream = streams[0];
    }
    if (bytes > cap[slot]) {
      if (p[slot]) CU(cudaFree(p[slot]));
      p[slot] = nullptr; cap[slot] = 0;
      size_t want = bytes + bytes / 8 + 256;
      CU(cudaMalloc(&p[slot], want));
      cap[slot] = want;
    }
    *out = p[slot];
    return HB_OK;
  }
};
thread_local Scratch g_scratch;

// Resident CTA slots of a kernel on the current device (SMs x occupancy), cached per (device, function).
// Per-kernel caches, keyed by the kernel handle.  Handles of NVRTC-compiled systems are reused by the driver after
// hb_system_free (cudaLibraryUnload), so freeing a system purges its entries (forget_kernel) — a stale "opt-in shared memory
// already granted" entry makes the next kernel at that address fail to launch.
struct ShapeOcc { int per_sm[17]; };
std::mutex g_kcache_mu;
std::map<std::tuple<int, const void*, int, size_t>, int> g_occ_cache;                   // (device, kernel, block, dyn smem) -> resident CTAs
std::map<std::pair<int, const void*>, size_t> g_granted_dyn;                            // opt-in dynamic shared memory granted per kernel
std::map<std::tuple<int, const void*, int, int, int>, ShapeOcc> g_shape_cache;          // pick_shape: CTAs per SM per candidate size
void forget_kernel(const void* fn) {
  std::lock_guard<std::mutex> lk(g_kcache_mu);
  for (auto it = g_occ_cache.begin(); it != g_occ_cache.end();) it = std::get<1>(it->first) == fn ? g_occ_cache.erase(it) : std::next(it);
  for (auto it = g_granted_dyn.begin(); it != g_granted_dyn.end();) it = it->first.second == fn ? g_granted_dyn.erase(it) : std::next(it);
  for (auto it = g_shape_cache.begin(); it != g_shape_cache.end();) it = std::get<1>(it->first) == fn ? g_shape_cache.erase(it) : std::next(it);
}

int resident_ctas(const void* fn, int block, size_t dyn_smem) {
  auto& cache = g_occ_cache;
  auto& max_dyn = g_granted_dyn;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  std::lock_guard<std::mutex> lk(g_kcache_mu);
  auto it = cache.find(std::make_tuple(dev, fn, block, dyn_smem));
  if (it != cache.end()) return it->second;
  int sms = 0, per_sm = 0;
  size_t& granted = max_dyn[std::make_pair(dev, fn)];
  if (dyn_smem > 48 * 1024 && dyn_smem > granted) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem) != cudaSuccess) { cudaGetLastError(); return -1; }
    granted = dyn_smem;
  }
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, dyn_smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  const int slots = sms * (per_sm > 0 ? per_sm : 1);
  cache[std::make_tuple(dev, fn, block, dyn_smem)] = slots;
  return slots;
}

// Launch policy.  Batches that fill the chip run as ONE resident wave (grid = SMs x occupancy, 148 x k on B200): every CTA
// stages the sin/cos table once (one bulk copy) and its warps walk tiles of 32 trajectories, b + G (w + W r) — every round
// of tiles spreads over all CTAs, so the last partial round is balanced over the SMs (engine/hb_engine.cuh HB_KERNEL_BODY).
// Every launch carries the programmatic-stream-serialization attribute: the kernels call griddepcontrol.launch_dependents
// first thing and griddepcontrol.wait before their first global read, so in a stream (or captured graph) of back-to-back
// steps the next kernel's launch latency, table staging and first L2 prefetches hide under this kernel's tail.
// The kernels index trajectories with 32 bits: callers (run_batch) hand at most HB_MAX_LAUNCH_N trajectories per launch.
hb_status launch(const void* fn, const HbKArgs& a, long long grid, cudaStream_t st, int block, size_t dyn_smem) {
  if (grid <= 0) return HB_OK;
  static const bool pdl = std::getenv("HB_NO_PDL") == nullptr;
  if (resident_ctas(fn, block, dyn_smem) < 0) return fail(HB_ERR_CUDA, "kernel needs more dynamic shared memory than the device offers");
  if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;
  void* args[] = {(void*)&a};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = dyn_smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t le = cudaLaunchKernelExC(&cfg, fn, args);
  if (le != cudaSuccess) {   // say what was asked for: a launch failure is otherwise undiagnosable from the status alone
    cudaGetLastError();
    cudaFuncAttributes fa;
    char buf[384];
    if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess)
      std::snprintf(buf, sizeof buf, " [grid %lld, block %d, dynamic smem %zu; kernel: %d registers, %zu B static smem, max %d threads, max dynamic smem %d]",
                    grid, block, dyn_smem, fa.numRegs, fa.sharedSizeBytes, fa.maxThreadsPerBlock, fa.maxDynamicSharedSizeBytes);
    else { cudaGetLastError(); std::snprintf(buf, sizeof buf, " [grid %lld, block %d, dynamic smem %zu]", grid, block, dyn_smem); }
    return fail(HB_ERR_CUDA, std::string("cudaLaunchKernelExC: ") + cudaGetErrorString(le) + buf);
  }
  return HB_OK;
}

// Dynamic shared memory of one launch; mirrors the layout documented at HB_DYN_DOUBLES in engine/hb_engine.cuh.
bool stepping_kernel(int kid) { return kid == K_STEP_RK4 || kid == K_STEP_RKF45 || kid == K_EVOLVE_RK4 || kid == K_EVOLVE_RKF45; }
size_t dyn_smem_bytes(const hb_system* s, int kid, int block, int in_d, int out_d, int kernel_layout) {
  size_t b;
  if (s->n >= HB_BIG_N) b = (size_t)s->dyn_doubles * sizeof(double) * block;
  else {
    b = s->trig ? (size_t)HB_TAB_BYTES : 0;
    if (s->heavy && stepping_kernel(kid) && in_d % 2 == 0) b += (size_t)2 * in_d * sizeof(double) * block;   // cp.async stage (engine: ASYNC)
  }
  if (kernel_layout == 2 && out_d % 2 == 0 && out_d <= HB_WSTORE_MAXD) b += (size_t)out_d * sizeof(double) * block;
  // Empirical: a kernel that asks for NO shared memory at all gets the whole 256 KB as L1, and the HBM-bound two-body step runs
  // 13 % slower that way than with a 64 KB carve-out (25.4 vs 22.1 us, profiles/r2q/ab_two_body.txt; pendulum and 1-D spring
  // within 2 %).  Every launch therefore asks for at least 8 KB per CTA; the kernels never touch the padding.
  if (b < 8192) b = 8192;
  static const size_t extra = [] { const char* e = std::getenv("HB_EXTRA_SMEM"); long t = e ? std::atol(e) : 0; return (t > 0 && t <= 200000) ? (size_t)t : (size_t)0; }();
  return b + extra;   // HB_EXTRA_SMEM: experiment knob (occupancy / L1 carve-out studies)
}
// Launch shape of one kernel of a system: CTA size and grid.
// Large systems: HB_BLOCK_OF threads (their shared-memory layout is compiled for it), one resident wave.
// Small systems: the kernels are tile-scheduled (engine/hb_engine.cuh HB_KERNEL_BODY: warp w of CTA b walks tiles
// b + G (w + W r)), so a launch is R = tiles / (resident warps) rounds long and its LAST round is only frac(R) full —
// the few warps of a nearly empty last round run alone at latency-bound speed while the rest of the SM idles.  Measured on
// the double pendulum, 1,048,576 trajectories, one RK4 step per launch (profiles/r2c/ab_double_pendulum.txt):
//   5 CTAs x 128 threads / SM  (R = 11.07)  19.6 us      1 CTA x 512 (R = 13.84)  16.9 us
//   2 CTAs x 384               (R =  9.23)  18.8 us      1 CTA x 768 (R =  9.23)  17.5 us
// i.e. what counts is (1) few CTAs per SM — every CTA stages its own 32 KB sin/cos table image per launch — and (2) a
// last round that is nearly full; beyond 4 warps per scheduler more resident warps buy nothing (the step is issue-bound).
// pick_shape therefore scores every CTA size of 4..16 warps the kernel can run (occupancy queried once per kernel and
// cached) with   waste = [frac > 0] max(0, RHO - frac x warps/SM)  +  TABLE x CTAs/SM   (in units of one tile's issue time;
// RHO = 8: one warp alone needs ~8x its issue-bound share, TABLE = 2) and takes the cheapest, ties to more warps.
// HB_BLOCK / HB_GRID_WAVES override (experiments).
struct LaunchShape { int block; long long grid; int contiguous = 0; };
LaunchShape pick_shape(const hb_system* s, int kid, const void* fn, long long n_traj, int in_d, int out_d, int kernel_layout, bool heavy) {
  static const int block_env = [] { const char* e = std::getenv("HB_BLOCK"); int t = e ? std::atoi(e) : 0; return (t >= 32 && t <= 1024 && t % 32 == 0) ? t : 0; }();
  static const double waves_env = [] { const char* e = std::getenv("HB_GRID_WAVES"); double t = e ? std::atof(e) : 0.0; return (t > 0 && t <= 4096) ? t : 0.0; }();
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { cudaGetLastError(); sms = 148; }
  auto one_wave = [&](int block) -> LaunchShape {
    const int slots = resident_ctas(fn, block, dyn_smem_bytes(s, kid, block, in_d, out_d, kernel_layout));
    long long blocks = (n_traj + block - 1) / block;
    const long long cap = (long long)((double)(slots > 0 ? slots : sms) * (waves_env > 0 ? waves_env : 1.0));
    if (blocks > cap) blocks = cap;
    return LaunchShape{block, blocks < 1 ? 1 : blocks};
  };
  // Large systems: long kernels (150 us for the 12-link chain) on 2 CTAs per SM — TWO waves with the contiguous tile map (the
  // CTAs of the second wave go to whichever SM frees up first: dynamic balance, 6 % against one static wave, profiles/r2n)
  if (s->n >= HB_BIG_N) { LaunchShape sh = one_wave(HB_BLOCK_OF(s->n)); const long long need = (n_traj + sh.block - 1) / sh.block; sh.grid = std::min<long long>(need, 2 * sh.grid); sh.contiguous = 1; return sh; }
  if (block_env) return one_wave(block_env);
  // HBM-bound launches (light systems, the one-evaluation kernels): what counts is bytes in flight and how often the 32 KB
  // table image is staged (once per CTA).  Measured per system over CTA sizes 128/256/512 x 1/2/4 waves x both tile maps
  // (profiles/r2x, r2y; round-2 rule before that: 128 threads, four waves, profiles/r2s):
  //   one-evaluation kernels (hamEqs, toPhase, ...): 512-thread CTAs with the contiguous tile map everywhere; ONE wave for the
  //     small table-staging systems (pendulum hamEqs 14.2 -> 10.5 us = 0.98 of the measured HBM peak, double pendulum 13.7 ->
  //     12.3, room 14.3 -> 11.3), FOUR waves for records of >= 6 doubles and for systems without a table (triple pendulum
  //     22.0 -> 19.4, two-body 25.3 -> 24.0, 1-D spring 10.9 -> 10.4);
  //   stepping kernels: a light system that stages the table (pendulum, 1.4) runs best as one spread wave of 512-thread
  //     CTAs (14.6 -> 13.6 us); the most HBM-bound ones without a table (two-body, 1-D spring: estimate below 1.2) as four
  //     contiguous waves of 128-thread CTAs (two-body 23.1 us against 25.1 for one spread wave, 1-D spring 10.9 against 11.4).
  if (!heavy) {
    const bool one_eval = !stepping_kernel(kid);
    const bool big_cta = one_eval || s->trig;
    LaunchShape sh = one_wave(big_cta ? 512 : 128);
    if (waves_env <= 0) {
      const long long need = (n_traj + sh.block - 1) / sh.block;
      if (one_eval) {
        const int waves = (s->trig && in_d < 6) ? 1 : 4;
        sh.grid = std::min<long long>(need, waves * sh.grid);
        sh.contiguous = 1;
      } else if (!s->trig && s->intensity < 1.2) {
        sh.grid = std::min<long long>(need, 4 * sh.grid);
        sh.contiguous = 1;
      }
    }
    return sh;
  }
  // CTAs per SM for every candidate size, cached per kernel
  typedef ShapeOcc Occ;
  std::mutex& mu = g_kcache_mu;
  auto& cache = g_shape_cache;
  const auto key = std::make_tuple(dev, fn, in_d, out_d, kernel_layout == 2 ? 1 : 0);
  Occ occ;
  bool have = false;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { occ = it->second; have = true; }
  }
  if (!have) {
    cudaFuncAttributes fa;
    int max_threads = 128;
    if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess) max_threads = fa.maxThreadsPerBlock; else cudaGetLastError();
    for (int w = 0; w <= 16; w++) {
      occ.per_sm[w] = 0;
      if (w < 4 || 32 * w > max_threads) continue;
      const int slots = resident_ctas(fn, 32 * w, dyn_smem_bytes(s, kid, 32 * w, in_d, out_d, kernel_layout));
      occ.per_sm[w] = slots > 0 ? slots / sms : 0;
    }
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = occ;
  }
  const double tiles = (double)((n_traj + 31) / 32);
  const double RHO = 8.0, TABLE = 2.0;
  int best_w = 4, best_c = 1, best_wsm = 0;
  double best_cost = 1e300;
  for (int w = 4; w <= 16; w++) {
    for (int c = 1; c <= occ.per_sm[w]; c++) {         // c CTAs of w warps per SM
      const int wsm = w * c;
      if (wsm > 24) break;
      const double R = tiles / ((double)sms * wsm);
      const double frac = R - std::floor(R);
      double cost = (frac > 1e-9 ? std::max(0.0, RHO - frac * wsm) : 0.0) + TABLE * c;
      if (wsm < 12 && tiles >= (double)sms * 12) cost += 2.0 * (12 - wsm);     // too few warps to cover the FP64 latency
      if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && wsm > best_wsm)) { best_cost = cost; best_w = w; best_c = c; best_wsm = wsm; }
    }
  }
  if (best_wsm == 0) return one_wave(128);
  const int block = 32 * best_w;
  long long blocks = (n_traj + block - 1) / block;
  const long long cap = (long long)((double)sms * best_c * (waves_env > 0 ? waves_env : 1.0));
  if (blocks > cap) blocks = cap;
  return LaunchShape{block, blocks < 1 ? 1 : blocks};
}
// heavy: is this launch issue-bound?  Stepping kernels of HEAVY systems, and any stepping kernel that takes several steps (or
// adaptive solves: ~25 RHS each) per trajectory.
bool heavy_launch(const hb_system* s, int kid, const HbKArgs& a) {
  if (kid == K_STEP_RK4) return s->heavy || a.nsteps >= 4;
  return kid == K_STEP_RKF45 || kid == K_EVOLVE_RK4 || kid == K_EVOLVE_RKF45;
}
hb_status launch_sys(const hb_system* s, int kid, const void* fn, const HbKArgs& a, long long n_traj, cudaStream_t st, int in_d, int out_d) {
  const LaunchShape sh = pick_shape(s, kid, fn, n_traj, in_d, out_d, a.layout, heavy_launch(s, kid, a));
  HbKArgs ac = a;
  static const int contig_env = [] { const char* e = std::getenv("HB_CONTIGUOUS"); return e ? std::atoi(e) : -1; }();   // experiment knob
  ac.contiguous = contig_env >= 0 ? contig_env : sh.contiguous;
  return launch(fn, ac, sh.grid, st, sh.block, dyn_smem_bytes(s, kid, sh.block, in_d, out_d, a.layout));
}

void fill_params(const hb_system* s, HbKArgs& a) {
  std::memset(&a, 0, sizeof a);
  for (size_t k = 0; k < s->params.size() && k < HB_MAXP; k++) a.prm[k] = s->params[k];
}

// Generic batch runner.  in_d / out_d = doubles per trajectory on each side; `out_batches` output
// batches are stored back to back (evolve).  `ts` (host, s doubles) is uploaded when non-null.
hb_status run_batch(const hb_system* sys, int kid, HbKArgs a, int64_t N, hb_memspace mem, const double* in, int in_d,
                    double* out, int out_d, int out_batches, int32_t* flags, const double* ts, int s, void* stream) {
  if (!sys) return fail(HB_ERR_INVALID, "null system");
  if (N < 0) return fail(HB_ERR_INVALID, "negative batch size");
  if (N == 0) return HB_OK;
  if (!in || !out) return fail(HB_ERR_INVALID, "null batch pointer");
  // array-of-records pointers are cast to double2 by the kernels: a misaligned pointer would be a sticky device fault
  if (a.layout == HB_LAYOUT_AOS && (((uintptr_t)in | (uintptr_t)out) & 15) != 0 && in_d % 2 == 0 && out_d % 2 == 0)
    return fail(HB_ERR_INVALID, "array-of-Phases batches must be 16-byte aligned");
  if ((((uintptr_t)in | (uintptr_t)out) & 7) != 0 || ((uintptr_t)flags & 3) != 0) return fail(HB_ERR_INVALID, "misaligned batch pointer");
  if (N > HB_MAX_LAUNCH_N) {   // the kernels index trajectories with 32 bits: larger batches go in slices
    if (a.layout != HB_LAYOUT_AOS || out_batches != 1)
      return fail(HB_ERR_UNSUPPORTED, "more than 2^31 - 2^26 trajectories per call need the array-of-Phases layout (and evolve calls must be sliced by the caller)");
    for (int64_t i0 = 0; i0 < N; i0 += HB_MAX_LAUNCH_N) {
      const int64_t n = N - i0 < HB_MAX_LAUNCH_N ? N - i0 : HB_MAX_LAUNCH_N;
      hb_status rs = run_batch(sys, kid, a, n, mem, in + (size_t)i0 * in_d, in_d, out + (size_t)i0 * out_d, out_d, 1, flags ? flags + i0 : nullptr, ts, s, stream);
      if (rs) return rs;
    }
    return HB_OK;
  }
  hb_status rc = need_device();
  if (rc) return rc;
  const void* fn = nullptr;
  rc = const_cast<hb_system*>(sys)->kernel(kid, &fn);
  if (rc) return rc;
  a.N = N;
  const size_t in_bytes = (size_t)N * in_d * sizeof(double), out_bytes = (size_t)N * out_d * out_batches * sizeof(double);
  if (mem == HB_MEM_DEVICE) {
    cudaStream_t st = (cudaStream_t)stream;
    a.in = in; a.out = out; a.flags = flags;
    double* dts = nullptr;
    if (ts) {
      CU(cudaMallocAsync((void**)&dts, sizeof(double) * s, st));
      CU(cudaMemcpyAsync(dts, ts, sizeof(double) * s, cudaMemcpyHostToDevice, st));   // ts is tiny; pageable copy is staged by the driver
      a.ts = dts;
    }
    rc = launch_sys(sys, kid, fn, a, N, st, in_d, out_d);
    if (dts) cudaFreeAsync(dts, st);
    return rc;
  }
  // HB_MEM_HOST, page-locked (mapped) caller buffers: ONE kernel reads the Phases straight out of host memory and writes
  // the results straight back over PCIe — no staging copies, no fill/drain of a chunk pipeline, reads and writes overlap
  // inside the kernel.  SM loads from host memory run at the link rate (50.9 GB/s measured, profiles/r1h/zc.txt); stores
  // do only when every warp instruction writes whole sectors (53.6 vs 12.7 GB/s), hence the warp-transposed store
  // (kernel layout 2, records of <= HB_WSTORE_MAXD doubles) or the SOA layout.  HB_HOST_DIRECT=0 forces staging.
  static const int direct_mode = [] { const char* e = std::getenv("HB_HOST_DIRECT"); return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }();
  const bool direct_shape = a.layout == HB_LAYOUT_SOA || (out_d % 2 == 0 && out_d <= HB_WSTORE_MAXD);
  auto pinned_dev = [](const void* h, void** d) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
    *d = at.devicePointer;
    return true;
  };
  void *din_h = nullptr, *dout_h = nullptr, *dfl_h = nullptr;
  const bool mapped = direct_mode && direct_shape && pinned_dev(in, &din_h) && pinned_dev(out, &dout_h) && (!flags || pinned_dev(flags, &dfl_h));
  if (mapped && direct_mode == 1) {
    {
      void* dts = nullptr;
      if ((rc = g_scratch.get(0, ts ? sizeof(double) * s : 8, &dts))) return rc;
      cudaStream_t st = g_scratch.stream;
      if (ts) { CU(cudaMemcpyAsync(dts, ts, sizeof(double) * s, cudaMemcpyHostToDevice, st)); a.ts = (const double*)dts; }
      a.in = (const double*)din_h; a.out = (double*)dout_h; a.flags = (int*)dfl_h;
      if (a.layout == HB_LAYOUT_AOS) a.layout = 2;   // warp-transposed stores (engine/hb_engine.cuh hb_store)
      a.host_io = 1;
      if ((rc = launch_sys(sys, kid, fn, a, N, st, in_d, out_d))) return rc;
      CU(cudaStreamSynchronize(st));
      return HB_OK;
    }
  }
  // HB_MEM_HOST: stage through per-thread device scratch and wait.  Large AOS batches are cut into chunks that
  // flow through three streams, so the H2D copy of chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1
  // overlap (PCIe is full duplex; the copy engines and the SMs run concurrently).
  void *din = nullptr, *dout = nullptr, *dfl = nullptr;
  const bool inplace = (const void*)in == (const void*)out && in_bytes == out_bytes;
  if ((rc = g_scratch.get(0, in_bytes + (ts ? sizeof(double) * s : 0), &din))) return rc;
  if (inplace) dout = din; else if ((rc = g_scratch.get(1, out_bytes, &dout))) return rc;
  if (flags && (rc = g_scratch.get(2, sizeof(int32_t) * N, &dfl))) return rc;
  // Three decoupled in-order queues joined by events: every H2D copy sits back to back on the upload stream, the kernel
  // of chunk c waits only for its own upload, the D2H copy of chunk c only for its own kernel — both PCIe directions
  // stay busy for the whole call except for one chunk of fill and one of drain: T ~ (1 + 1/chunks) x one-way copy time
  // + chunks x per-chunk submission cost.
  static const int chunks_env = [] { const char* e = std::getenv("HB_HOST_CHUNKS"); int t = e ? std::atoi(e) : 0; return (t >= 1 && t <= 64) ? t : 0; }();
  static const bool graph_ok = [] { const char* e = std::getenv("HB_HOST_GRAPH"); return !(e && e[0] == '0'); }();
  // Page-locked caller buffers: the whole chunk pipeline is captured ONCE into a CUDA graph, cached per (kernel, buffers,
  // arguments), and replayed with a single cudaGraphLaunch on later calls — a stepping loop over fixed buffers then pays
  // one API call per step instead of 7 per chunk, which is what allows small chunks (short fill/drain).
  auto pinned = [](const void* h) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  };
  const bool chunkable = a.layout == HB_LAYOUT_AOS && out_batches == 1 && !ts && N >= (1 << 16);
  // HB_HOST_DIRECT=2 (hybrid): the copy engine uploads the chunks, the kernel of each chunk stores straight into the
  // caller's page-locked output (warp-transposed) — no download copies, no duplex copy-engine traffic.
  const bool hybrid = mapped && direct_mode == 2 && chunkable;
  const bool use_graph = graph_ok && chunkable && pinned(in) && pinned(out) && (!flags || pinned(flags));
  int64_t chunks = 1;
  if (chunkable) {
    // >= 2 MiB per chunk when replayed from a graph, >= 8 MiB when every chunk costs 7 stream API calls
    const size_t per = use_graph ? ((size_t)2 << 20) : ((size_t)8 << 20);
    chunks = chunks_env ? chunks_env : (int64_t)(in_bytes / per);
    if (!chunks_env && chunks > (use_graph ? 16 : 4)) chunks = use_graph ? 16 : 4;
    if (chunks < 1) chunks = 1;
  }
  cudaStream_t s_up = g_scratch.streams[0], s_k = g_scratch.streams[1], s_down = g_scratch.streams[2];
  if ((rc = g_scratch.need_events((int)(2 * chunks) + 1))) return rc;
  (void)pick_shape(sys, kid, fn, N, in_d, out_d, hybrid ? 2 : a.layout, heavy_launch(sys, kid, a));   // (keeps the occupancy queries out of a capture)
  (void)pick_shape(sys, kid, fn, (N + chunks - 1) / chunks, in_d, out_d, hybrid ? 2 : a.layout, heavy_launch(sys, kid, a));
  auto enqueue = [&]() -> hb_status {
    if (ts) {
      double* dts = (double*)((char*)din + in_bytes);
      CU(cudaMemcpyAsync(dts, ts, sizeof(double) * s, cudaMemcpyHostToDevice, s_up));
      a.ts = dts;
    }
    for (int64_t c = 0; c < chunks; c++) {
      const int64_t i0 = N * c / chunks, i1 = N * (c + 1) / chunks, n = i1 - i0;
      const double* hin = in + (size_t)i0 * in_d;
      double* hout = out + (size_t)i0 * out_d;
      double* cin = (double*)din + (size_t)i0 * in_d;
      double* cout = inplace ? cin : (double*)dout + (size_t)i0 * out_d;
      cudaEvent_t up = g_scratch.events[2 * c], done = g_scratch.events[2 * c + 1];
      if (chunks == 1) {   // whole batch (any layout, evolve output): copy everything
        CU(cudaMemcpyAsync(din, in, in_bytes, cudaMemcpyHostToDevice, s_up));
        if (flags) CU(cudaMemcpyAsync(dfl, flags, sizeof(int32_t) * N, cudaMemcpyHostToDevice, s_up));
      } else {
        CU(cudaMemcpyAsync(cin, hin, (size_t)n * in_d * sizeof(double), cudaMemcpyHostToDevice, s_up));
        if (flags && !hybrid) CU(cudaMemcpyAsync((int32_t*)dfl + i0, flags + i0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s_up));
      }
      CU(cudaEventRecord(up, s_up));
      CU(cudaStreamWaitEvent(s_k, up, 0));
      HbKArgs ac = a;
      if (chunks == 1) { ac.N = N; ac.in = (const double*)din; ac.out = (double*)dout; ac.flags = flags ? (int*)dfl : nullptr; }
      else { ac.N = n; ac.in = cin; ac.out = cout; ac.flags = flags ? (int*)dfl + i0 : nullptr; }
      if (hybrid) {
        ac.out = (double*)dout_h + (size_t)i0 * out_d;
        ac.flags = flags ? (int*)dfl_h + i0 : nullptr;   // read-modify-write in host memory, only for flagged trajectories
        ac.layout = 2;
      }
      hb_status lrc = launch_sys(sys, kid, fn, ac, chunks == 1 ? N : n, s_k, in_d, out_d);
      if (lrc) return lrc;
      if (hybrid) { if (c + 1 == chunks) { CU(cudaEventRecord(done, s_k)); CU(cudaStreamWaitEvent(s_down, done, 0)); } continue; }
      CU(cudaEventRecord(done, s_k));
      CU(cudaStreamWaitEvent(s_down, done, 0));
      if (chunks == 1) {
        CU(cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, s_down));
        if (flags) CU(cudaMemcpyAsync(flags, dfl, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, s_down));
      } else {
        CU(cudaMemcpyAsync(hout, cout, (size_t)n * out_d * sizeof(double), cudaMemcpyDeviceToHost, s_down));
        if (flags) CU(cudaMemcpyAsync(flags + i0, (int32_t*)dfl + i0, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s_down));
      }
    }
    return HB_OK;
  };
  if (use_graph && chunks > 1) {
    HostPipeKey key;
    std::memset(&key, 0, sizeof key);
    key.fn = fn; key.in = in; key.out = out; key.flags = flags; key.din = din; key.dout = dout; key.dfl = dfl;
    key.N = N; key.chunks = chunks; key.in_d = in_d; key.out_d = out_d; key.dyn = sys->dyn_doubles; key.pad_ = hybrid ? 1 : 0; key.a = a;
    cudaGraphExec_t exec = g_scratch.find_pipe(key);
    if (!exec) {
      cudaGraph_t graph = nullptr;
      CU(cudaStreamBeginCapture(s_up, cudaStreamCaptureModeThreadLocal));
      rc = enqueue();
      cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
      if (!rc) {   // join the kernel and download queues back into the origin stream
        cudaEvent_t fin = g_scratch.events[2 * chunks];
        e1 = cudaEventRecord(fin, s_down);
        if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(s_up, fin, 0);
      }
      e2 = cudaStreamEndCapture(s_up, &graph);
      if (rc || e1 != cudaSuccess || e2 != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc) return rc;
        exec = nullptr;   // capture not possible here: fall through to plain stream submission
      } else {
        cudaError_t e3 = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e3 != cudaSuccess) { cudaGetLastError(); exec = nullptr; }
        else g_scratch.add_pipe(key, exec);
      }
    }
    if (exec) {
      CU(cudaGraphLaunch(exec, s_up));
      CU(cudaStreamSynchronize(s_up));
      return HB_OK;
    }
  }
  if ((rc = enqueue())) return rc;
  CU(cudaStreamSynchronize(s_down));   // the last download is the last operation of the call
  return HB_OK;
}

bool bad_layout(hb_layout l) { return l != HB_LAYOUT_AOS && l != HB_LAYOUT_SOA; }

}  // namespace

// ================================================================================ C ABI ======
extern "C" {

// internal (ensemble.cpp): sets the calling thread's error message
hb_status hb_internal_fail(hb_status s, const char* msg) { return fail(s, msg ? msg : ""); }

int32_t hb_abi_version(void) { return HB_ABI_VERSION; }
const char* hb_last_error(void) { return g_err.c_str(); }

hb_status hb_device_count(int32_t* count) {
  if (!count) return fail(HB_ERR_INVALID, "null count");
  *count = 0;
  hb_status rc = need_device();
  if (rc) return rc;
  int n = 0;
  CU(cudaGetDeviceCount(&n));
  *count = n;
  return HB_OK;
}
hb_status hb_set_device(int32_t device) {
  hb_status rc = need_device();
  if (rc) return rc;
  CU(cudaSetDevice(device));
  return HB_OK;
}

hb_status hb_system_builtin(hb_builtin id, const double* params, int32_t n_params, hb_system** out) {
  if (!out) return fail(HB_ERR_INVALID, "null out");
  *out = nullptr;
  if ((int)id < 0 || (int)id >= HB_SYS__COUNT) return fail(HB_ERR_INVALID, "unknown builtin system id");
  if (n_params < 0 || (n_params > 0 && !params)) return fail(HB_ERR_INVALID, "bad params");
  hb::SystemSpec spec;
  if (!hb::builtin_spec(id, spec)) return fail(HB_ERR_INVALID, "no such builtin");
  hb_system* s = new hb_system();
  s->m = spec.m; s->n = spec.n; s->builtin = (int)id;
  hb::builtin_params(id, params, n_params, s->params);
  std::vector<double> dflt;
  hb::builtin_params(id, nullptr, 0, dflt);
  s->baked = std::getenv("HB_NO_BAKED") == nullptr && dflt == s->params;
  s->dyn_doubles = hb_aot_dyn_doubles((int)id + (s->baked ? HB_SYS__COUNT : 0));
  if (s->baked) spec.baked_params = dflt;
  if (hb_aot_kargs_size() != sizeof(HbKArgs)) { delete s; return fail(HB_ERR_INVALID, "internal: host/device HbKArgs layout mismatch"); }
  hb::GeneratedSystem g;
  std::string err;
  if (hb::generate_system(spec, std::string("HbSys_") + hb::builtin_name(id) + (s->baked ? "_dflt" : ""), g, err)) { s->source = g.source; s->rhs_cost = g.rhs_cost; s->heavy = g.heavy; s->trig = g.trig; s->intensity = g.intensity; }
  *out = s;
  return HB_OK;
}

hb_status hb_system_from_tape(int32_t m, int32_t n, const double* inertia, const hb_tape* f, const hb_tape* u,
                              int32_t u_on_cartesian, const double* params, int32_t n_params, hb_system** out) {
  if (!out) return fail(HB_ERR_INVALID, "null out");
  *out = nullptr;
  if (!inertia || !f || !u) return fail(HB_ERR_INVALID, "null argument");
  if (n < 1 || n > HB_MAX_N || m < 1 || m > HB_MAX_M) return fail(HB_ERR_INVALID, "dimensions out of range (1 <= n <= 16, 1 <= m <= 48)");
  if (n_params < 0 || n_params > 32 || (n_params > 0 && !params)) return fail(HB_ERR_INVALID, "bad params (at most 32)");
  if (f->n_in != n || f->n_out != m || !f->ops || !f->outs || f->n_ops < 1) return fail(HB_ERR_TAPE, "f tape must map n inputs to m outputs");
  if (u->n_in != (u_on_cartesian ? m : n) || u->n_out != 1 || !u->ops || !u->outs || u->n_ops < 1)
    return fail(HB_ERR_TAPE, "u tape must map n (or m when u_on_cartesian) inputs to 1 output");
  hb::SystemSpec spec;
  spec.m = m; spec.n = n; spec.n_params = n_params; spec.u_on_cartesian = u_on_cartesian != 0;
  for (int i = 0; i < m; i++) { hb::InertiaTerm t; t.value = inertia[i]; spec.inertia.push_back(t); }
  spec.f_ops.assign(f->ops, f->ops + f->n_ops);
  spec.f_outs.assign(f->outs, f->outs + m);
  spec.u_ops.assign(u->ops, u->ops + u->n_ops);
  spec.u_out = u->outs[0];
  hb::GeneratedSystem g;
  std::string err;
  if (!hb::generate_system(spec, "HbSysJit", g, err)) return fail(HB_ERR_TAPE, err);
  hb_system* s = new hb_system();
  s->m = m; s->n = n; s->builtin = -1;
  s->params.assign(params, params + n_params);
  s->source = g.source;
  s->rhs_cost = g.rhs_cost;
  s->heavy = g.heavy;
  s->trig = g.trig;
  s->intensity = g.intensity;
  s->gen = g;
  s->dyn_doubles = HB_DYN_DOUBLES(n, g.ne);
  s->arch = jit_arch();
  s->no_code = std::getenv("HB_JIT_SKIP_COMPILE") != nullptr;   // diagnostics: symbolic stage only, system cannot launch
  const char* lz = std::getenv("HB_JIT_LAZY");
  s->lazy = lz ? std::atoi(lz) != 0 : (g.nj + g.nh > 64);
  if (!s->no_code && !s->lazy) {
    std::string log;
    if (!nvrtc_compile(hb::jit_translation_unit(g, "hbk"), s->arch, s->cubins[0], log)) { delete s; return fail(HB_ERR_COMPILE, log); }
  }
  *out = s;
  return HB_OK;
}

void hb_system_free(hb_system* sys) {
  if (!sys) return;
  for (const void* k : sys->jit_kernels) if (k) forget_kernel(k);   // the driver reuses kernel handles: drop what was cached under them
  for (auto& l : sys->libs) if (l) cudaLibraryUnload(l);
  delete sys;
}
hb_status hb_system_dims(const hb_system* sys, int32_t* m, int32_t* n) {
  if (!sys) return fail(HB_ERR_INVALID, "null system");
  if (m) *m = sys->m;
  if (n) *n = sys->n;
  return HB_OK;
}
size_t hb_system_source(const hb_system* sys, char* buf, size_t cap) {
  if (!sys) return 0;
  size_t need = sys->source.size() + 1;
  if (buf && cap) { size_t k = need < cap ? need : cap; std::memcpy(buf, sys->source.c_str(), k - 1); buf[k - 1] = 0; }
  return need;
}

int32_t hb_system_params(const hb_system* sys, double* buf, int32_t cap) {
  if (!sys) return 0;
  const int32_t n = (int32_t)sys->params.size();
  for (int32_t k = 0; buf && k < n && k < cap; k++) buf[k] = sys->params[k];
  return n;
}

hb_status hb_batch_ham_eqs(const hb_system* sys, int64_t N, hb_layout layout, hb_memspace mem, const double* y, double* dy,
                           int32_t* flags, void* stream) {
  if (!sys) return fail(HB_ERR_INVALID, "null system");
  if (bad_layout(layout)) return fail(HB_ERR_INVALID, "bad layout");
  HbKArgs a; fill_params(sys, a); a.layout = layout;
  return run_batch(sys, K_HAM_EQS, a, N, mem, y, 2 * sys->n, dy, 2 * sys->n, 1, flags, nullptr, 0, stream);
}

hb_status hb_batch_step(const hb_system* sys, hb_integrator integ, double dt, int32_t nsteps, int64_t N, hb_layout layout,
                        hb_memspace mem, const double* y_in, double* y_out, int32_t* flags, void* stream) {
  if (!sys) return fail(HB_ERR_INVALID, "null system");
  if (bad_layout(layout)) return fail(HB_ERR_INVALID, "bad layout");
  if (integ != HB_INTEG_RK4 && integ != HB_INTEG_RKF45_GSL) return fail(HB_ERR_INVALID, "unknown integrator");
  if (nsteps < 0) return fail(HB_ERR_INVALID, "negative nsteps");
  // RKF45_GSL with dt <= 0: the kernel returns every Phase unchanged, as `stepHam r` does for r <= 0 (hmatrix-gsl's
  // `while (t < t1)` loop never runs)
  HbKArgs a; fill_params(sys, a); a.layout = layout; a.dt = dt; a.dt6 = dt / 6.0; a.dth = 0.5 * dt; a.nsteps = nsteps;
  return run_batch(sys, integ == HB_INTEG_RK4 ? K_STEP_RK4 : K_STEP_RKF45, a, N, mem, y_in, 2 * sys->n, y_out, 2 * sys->n, 1, flags,
                   nullptr, 0, stream);
}

hb_status hb_batch_evolve(const hb_system* sys, hb_integrator integ, int32_t rk4_substeps, int64_t N, hb_layout layout,
                          hb_memspace mem, const double* y0, const double* ts, int32_t s, double* out, int32_t* flags, void* stream) {
  if (!sys) return fail(HB_ERR_INVALID, "null system");
  if (bad_layout(layout)) return fail(HB_ERR_INVALID, "bad layout");
  if (integ != HB_INTEG_RK4 && integ != HB_INTEG_RKF45_GSL) return fail(HB_ERR_INVALID, "unknown integrator");
  if (!ts || s < 2) return fail(HB_ERR_INVALID, "evolveHam needs at least two grid times (2 <= s, src/Numeric/Hamilton.hs:435)");
  for (int k = 1; k < s; k++) if (!(ts[k] >= ts[k - 1])) return fail(HB_ERR_INVALID, "time grid must be non-decreasing");
  if (integ == HB_INTEG_RK4 && rk4_substeps < 1) return fail(HB_ERR_INVALID, "rk4_substeps must be >= 1");
  // the reference's initial step is (ts[1] - ts[0]) / 100 (src/Numeric/Hamilton.hs:447): zero makes its solver spin forever
  if (integ == HB_INTEG_RKF45_GSL && !(ts[1] > ts[0])) return fail(HB_ERR_INVALID, "RKF45_GSL needs ts[1] > ts[0]: the initial step size is (ts[1] - ts[0]) / 100");
  HbKArgs a; fill_params(sys, a); a.layout = layout; a.s = s; a.substeps = rk4_substeps;
  return run_batch(sys, integ == HB_INTEG_RK4 ? K_EVOLVE_RK4 : K_EVOLVE_RKF45, a, N, mem, y0, 2 * sys->n, out, 2 * sys->n, s, flags, ts,
                   s, stream);
}

hb_status hb_batch_to_phase(const hb_system* sys, int64_t N, hb_layout layout, hb_memspace mem, const double* c, double* y, void* stream) {
  if (!sys) return fail(HB_ERR_INVALID, "null system");
  if (bad_layout(layout)) return fail(HB_ERR_INVALID, "bad layout");
  HbKArgs a; fill_params(sys, a); a.layout = layout;
  return run_batch(sys, K_TO_PHASE, a, N, mem, c, 2 * sys->n, y, 2 * sys->n, 1, nullptr, nullptr, 0, stream);
}
hb_status hb_batch_from_phase(const hb_system* sys, int64_t N, hb_layout layout, hb_memspace mem, const double* y, double* c,
                              int32_t* flags, void* stream) {
  if (!sys) return fail(HB_ERR_INVALID, "null system");
  if (bad_layout(layout)) return fail(HB_ERR_INVALID, "bad layout");
  HbKArgs a; fill_params(sys, a); a.layout = layout;
  return run_batch(sys, K_FROM_PHASE, a, N, mem, y, 2 * sys->n, c, 2 * sys->n, 1, flags, nullptr, 0, stream);
}
hb_status hb_batch_energies(const hb_system* sys, int64_t N, hb_layout layout, hb_memspace mem, const double* y, double* out4,
                            int32_t* flags, void* stream) {
  if (!sys) return fail(HB_ERR_INVALID, "null system");
  if (bad_layout(layout)) return fail(HB_ERR_INVALID, "bad layout");
  HbKArgs a; fill_params(sys, a); a.layout = layout;
  return run_batch(sys, K_ENERGIES, a, N, mem, y, 2 * sys->n, out4, 4, 1, flags, nullptr, 0, stream);
}
hb_status hb_batch_underlying_pos(const hb_system* sys, int64_t N, hb_layout layout, hb_memspace mem, const double* q, double* x,
                                  void* stream) {
  if (!sys) return fail(HB_ERR_INVALID, "null system");
  if (bad_layout(layout)) return fail(HB_ERR_INVALID, "bad layout");
  HbKArgs a; fill_params(sys, a); a.layout = layout;
  return run_batch(sys, K_UPOS, a, N, mem, q, sys->n, x, sys->m, 1, nullptr, nullptr, 0, stream);
}

hb_status hb_batch_init_random(const hb_system* sys, uint64_t seed, int64_t first, int64_t N, hb_layout layout, const double* lo,
                               const double* hi, double* y_device, void* stream) {
  if (!sys || !lo || !hi || !y_device) return fail(HB_ERR_INVALID, "null argument");
  if (bad_layout(layout) || N < 0 || first < 0) return fail(HB_ERR_INVALID, "bad argument");
  const int D = 2 * sys->n;
  if (2 * D > HB_MAXP) return fail(HB_ERR_INVALID, "too many components");
  hb_status rc = need_device();
  if (rc) return rc;
  HbKArgs a; std::memset(&a, 0, sizeof a);
  a.out = y_device; a.N = N; a.nsteps = D; a.layout = layout; a.seed = seed; a.first = first;
  for (int c = 0; c < D; c++) { a.prm[c] = lo[c]; a.prm[D + c] = hi[c]; }
  long long grid = (N * D + HB_BLOCK - 1) / HB_BLOCK;
  const int slots = resident_ctas(hb_aot_init_random(), HB_BLOCK, 0);
  if (slots > 0 && grid > 2LL * slots) grid = 2LL * slots;
  return launch(hb_aot_init_random(), a, grid, (cudaStream_t)stream, HB_BLOCK, 0);
}

// ---- single-trajectory mirrors ---------------------------------------------------------------
static hb_status one_flag(hb_status rc, int32_t flag) {
  if (rc) return rc;
  if (flag) {
    std::string m = "numerical failure:";
    if (flag & HB_FLAG_NOT_SPD) m += " mass matrix JtWJ not positive definite (reference: `inv` fails)";
    if (flag & HB_FLAG_NONFINITE) m += " non-finite state";
    if (flag & HB_FLAG_STEP_FAILED) m += " RKF45 step-size control failed";
    return fail(HB_ERR_NUMERIC, m);
  }
  return HB_OK;
}
#define NEED(sys, ...) do { if (!(sys)) return fail(HB_ERR_INVALID, "null system"); const void* ps_[] = {__VA_ARGS__}; \
  for (const void* p_ : ps_) if (!p_) return fail(HB_ERR_INVALID, "null pointer"); } while (0)

static void pack(const hb_system* s, const double* a, const double* b, double* y) {
  std::memcpy(y, a, sizeof(double) * s->n); std::memcpy(y + s->n, b, sizeof(double) * s->n);
}

hb_status hb_underlying_pos(const hb_system* sys, const double* q, double* x) {
  NEED(sys, q, x);
  return hb_batch_underlying_pos(sys, 1, HB_LAYOUT_AOS, HB_MEM_HOST, q, x, nullptr);
}
hb_status hb_pe(const hb_system* sys, const double* q, double* u) {
  NEED(sys, q, u);
  double y[2 * HB_MAX_N] = {0}, o[4]; int32_t fl = 0;
  std::memcpy(y, q, sizeof(double) * sys->n);   // p = 0: U does not depend on it
  hb_status rc = hb_batch_energies(sys, 1, HB_LAYOUT_AOS, HB_MEM_HOST, y, o, nullptr, nullptr);
  (void)fl;
  if (rc) return rc;
  *u = o[1];
  return HB_OK;
}
hb_status hb_momenta(const hb_system* sys, const double* q, const double* v, double* p) {
  NEED(sys, q, v, p);
  double c[2 * HB_MAX_N], y[2 * HB_MAX_N];
  pack(sys, q, v, c);
  hb_status rc = hb_batch_to_phase(sys, 1, HB_LAYOUT_AOS, HB_MEM_HOST, c, y, nullptr);
  if (rc) return rc;
  std::memcpy(p, y + sys->n, sizeof(double) * sys->n);
  return HB_OK;
}
hb_status hb_velocities(const hb_system* sys, const double* q, const double* p, double* v) {
  NEED(sys, q, p, v);
  double c[2 * HB_MAX_N], y[2 * HB_MAX_N]; int32_t fl = 0;
  pack(sys, q, p, y);
  hb_status rc = hb_batch_from_phase(sys, 1, HB_LAYOUT_AOS, HB_MEM_HOST, y, c, &fl, nullptr);
  rc = one_flag(rc, fl);
  if (rc) return rc;
  std::memcpy(v, c + sys->n, sizeof(double) * sys->n);
  return HB_OK;
}
static hb_status energies_p(const hb_system* sys, const double* q, const double* p, double* o4) {
  double y[2 * HB_MAX_N]; int32_t fl = 0;
  pack(sys, q, p, y);
  hb_status rc = hb_batch_energies(sys, 1, HB_LAYOUT_AOS, HB_MEM_HOST, y, o4, &fl, nullptr);
  return one_flag(rc, fl);
}
static hb_status energies_c(const hb_system* sys, const double* q, const double* v, double* o4) {
  double p[HB_MAX_N];
  hb_status rc = hb_momenta(sys, q, v, p);   // keC = (v . momenta) / 2 = keP at toPhase(c)
  if (rc) return rc;
  return energies_p(sys, q, p, o4);
}
hb_status hb_ke_p(const hb_system* sys, const double* q, const double* p, double* t) {
  NEED(sys, q, p, t); double o[4]; hb_status rc = energies_p(sys, q, p, o); if (!rc) *t = o[0]; return rc;
}
hb_status hb_hamiltonian(const hb_system* sys, const double* q, const double* p, double* h) {
  NEED(sys, q, p, h); double o[4]; hb_status rc = energies_p(sys, q, p, o); if (!rc) *h = o[2]; return rc;
}
hb_status hb_ke_c(const hb_system* sys, const double* q, const double* v, double* t) {
  NEED(sys, q, v, t); double o[4]; hb_status rc = energies_c(sys, q, v, o); if (!rc) *t = o[0]; return rc;
}
hb_status hb_lagrangian(const hb_system* sys, const double* q, const double* v, double* l) {
  NEED(sys, q, v, l); double o[4]; hb_status rc = energies_c(sys, q, v, o); if (!rc) *l = o[3]; return rc;
}
hb_status hb_ham_eqs(const hb_system* sys, const double* q, const double* p, double* dq, double* dp) {
  NEED(sys, q, p, dq, dp);
  double y[2 * HB_MAX_N], dy[2 * HB_MAX_N]; int32_t fl = 0;
  pack(sys, q, p, y);
  hb_status rc = hb_batch_ham_eqs(sys, 1, HB_LAYOUT_AOS, HB_MEM_HOST, y, dy, &fl, nullptr);
  rc = one_flag(rc, fl);
  if (rc) return rc;
  std::memcpy(dq, dy, sizeof(double) * sys->n); std::memcpy(dp, dy + sys->n, sizeof(double) * sys->n);
  return HB_OK;
}
hb_status hb_step_ham(const hb_system* sys, double r, const double* q, const double* p, double* q_out, double* p_out) {
  NEED(sys, q, p, q_out, p_out);
  double y[2 * HB_MAX_N], yo[2 * HB_MAX_N]; int32_t fl = 0;
  pack(sys, q, p, y);
  hb_status rc = hb_batch_step(sys, HB_INTEG_RKF45_GSL, r, 1, 1, HB_LAYOUT_AOS, HB_MEM_HOST, y, yo, &fl, nullptr);
  rc = one_flag(rc, fl);
  if (rc) return rc;
  std::memcpy(q_out, yo, sizeof(double) * sys->n); std::memcpy(p_out, yo + sys->n, sizeof(double) * sys->n);
  return HB_OK;
}
hb_status hb_evolve_ham(const hb_system* sys, const double* q0, const double* p0, const double* ts, int32_t s, double* out) {
  NEED(sys, q0, p0, ts, out);
  double y[2 * HB_MAX_N]; int32_t fl = 0;
  pack(sys, q0, p0, y);
  hb_status rc = hb_batch_evolve(sys, HB_INTEG_RKF45_GSL, 1, 1, HB_LAYOUT_AOS, HB_MEM_HOST, y, ts, s, out, &fl, nullptr);
  return one_flag(rc, fl);
}
hb_status hb_step_ham_c(const hb_system* sys, double r, const double* q, const double* v, double* q_out, double* v_out) {
  NEED(sys, q, v, q_out, v_out);
  double p[HB_MAX_N], qo[HB_MAX_N], po[HB_MAX_N];
  hb_status rc = hb_momenta(sys, q, v, p);                  // toPhase
  if (!rc) rc = hb_step_ham(sys, r, q, p, qo, po);          // stepHam
  if (!rc) rc = hb_velocities(sys, qo, po, v_out);          // fromPhase
  if (!rc) std::memcpy(q_out, qo, sizeof(double) * sys->n);
  return rc;
}
hb_status hb_evolve_ham_c(const hb_system* sys, const double* q0, const double* v0, const double* ts, int32_t s, double* out) {
  NEED(sys, q0, v0, ts, out);
  if (s < 2) return fail(HB_ERR_INVALID, "evolveHamC needs at least two grid times");
  double p[HB_MAX_N];
  hb_status rc = hb_momenta(sys, q0, v0, p);
  if (rc) return rc;
  std::vector<double> ph((size_t)s * 2 * sys->n);
  if ((rc = hb_evolve_ham(sys, q0, p, ts, s, ph.data()))) return rc;
  std::vector<int32_t> fl((size_t)s, 0);
  rc = hb_batch_from_phase(sys, s, HB_LAYOUT_AOS, HB_MEM_HOST, ph.data(), out, fl.data(), nullptr);   // fmap (fromPhase s)
  if (rc) return rc;
  int32_t any = 0; for (int32_t f : fl) any |= f;
  return one_flag(HB_OK, any);
}

}  // extern "C"
