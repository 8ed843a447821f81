// Compiles the DEVICE engine (hamilton_b200/csrc/engine/hb_engine.cuh, with HB_HOST_EMU) together with a generated `Sys`
// struct as HOST code: one emulated thread, CUDA built-ins stubbed.  Lets the CPU test-suite run the very code the
// kernels inline — hamEqs (both generated forms), the closed-form / LDL^T solves, classical RK4 (register and
// shared-memory variants), the GSL-RKF45 stepper + controller + evolve loop including its shortcut, the table-driven
// sincos and the Newton reciprocal — against the oracle.  Test infrastructure: the product never computes on the CPU.
#include <cmath>
#include <cstring>
#define HB_HOST_EMU 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__
#define __restrict__
#define __shared__ static
#define __align__(n)
#define __global__
#define __launch_bounds__(...)
#define __grid_constant__
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
struct HbEmuDim { unsigned x, y, z; };
static HbEmuDim threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {128, 1, 1}, gridDim = {1, 1, 1};
static inline int __double2hiint(double v) { long long b; std::memcpy(&b, &v, 8); return (int)(b >> 32); }
static inline int __double2loint(double v) { long long b; std::memcpy(&b, &v, 8); return (int)(b & 0xffffffffLL); }
static inline double __hiloint2double(int hi, int lo) { const unsigned long long b = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double v; std::memcpy(&v, &b, 8); return v; }
static inline unsigned __activemask() { return 1u; }
static inline void __syncwarp() {}
static inline void __syncthreads() {}
static inline unsigned long long __cvta_generic_to_shared(const void* p) { return (unsigned long long)p; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
// MUFU.RCP64H: the high word of an approximation of 1/d (relative error <= 2^-20, profiles/r1c/exp_rcp.txt); low word zero
static inline double hb_emu_rcp64h(double d) {
  double x = 1.0 / d;
  long long b; std::memcpy(&b, &x, 8);
  b &= ~0xffffffffLL;
  std::memcpy(&x, &b, 8);
  return x;
}
using std::exp; using std::log; using std::sqrt; using std::pow; using std::fabs; using std::tan; using std::atan2; using std::fma;
using std::asin; using std::acos; using std::atan; using std::sinh; using std::cosh; using std::tanh;
using std::asinh; using std::acosh; using std::atanh; using std::sin; using std::cos;
#include ENGINE_HEADER
#include SYS_SOURCE
typedef SYS_NAME S;
static constexpr int N = S::N, D = 2 * S::N;

static HbCtx ctx() { HbCtx cx; cx.tab_s = 0; hb_ctx_reset(cx); return cx; }
// HB_FLAG_* bits (integrator flags + deferred pivot test) | "left the fast domain" << 8
static int result(const HbCtx& cx, int flag) { return hb_ctx_flags(cx, flag) | (cx.oob ? 1 << 8 : 0); }   // (a failed pivot test alone also means "retry out of line": hb_retry)

extern "C" void dims(int* o) { o[0] = S::M; o[1] = S::N; o[2] = S::SYMH ? 1 : 0; o[3] = N >= HB_BIG_N ? 1 : 0; }

// hamEqs through the engine's own path (generated derivs/hpre/hpost + hb_spd_solve); returns flag | oob << 8
extern "C" int ham_eqs(const double* prm, const double* y, double* dy) {
  double w[S::M];
  S::inertia(prm, w);
  HbCtx cx = ctx();
  int flag = 0;
  hb_rhs<S, true>(cx, prm, w, y, dy, flag);
  return result(cx, flag);
}
extern "C" int rk4_steps(const double* prm, double* y_io, double dt, int nsteps) {
  double w[S::M];
  S::inertia(prm, w);
  HbCtx cx = ctx();
  int flag = 0;
  if constexpr (D >= 16) {   // the kernels' large-system path: RK vectors in (emulated) shared memory
    constexpr int B = HB_BLOCK_OF(S::N);
    double* sm = hb_rk_smem<D>();
    for (int c = 0; c < D; c++) sm[c * B] = y_io[c];
    for (int s = 0; s < nsteps; s++) hb_rk4_step_sm<S, true>(cx, prm, w, sm, dt, dt / 6.0, flag);
    for (int c = 0; c < D; c++) y_io[c] = sm[c * B];
  } else {
    double y[D];
    for (int c = 0; c < D; c++) y[c] = y_io[c];
    for (int s = 0; s < nsteps; s++) hb_rk4_step<S, true>(cx, prm, w, y, dt, dt / 6.0, 0.5 * dt, flag);
    for (int c = 0; c < D; c++) y_io[c] = y[c];
  }
  return result(cx, flag);
}
// stepHam iterated: a fresh GSL-RKF45 solve over (0, dt) per step, exactly as hb_traj_step_rkf45 does
extern "C" int rkf45_steps(const double* prm, double* y_io, double dt, int nsteps) {
  double w[S::M];
  S::inertia(prm, w);
  HbCtx cx = ctx();
  int flag = 0;
  double y[D];
  for (int c = 0; c < D; c++) y[c] = y_io[c];
  for (int s = 0; s < nsteps; s++) {
    HbEvolve<D> e;
    e.h = dt / 100;
    e.primed = false;
    double t = 0.0;
    hb_rkf45_to<S, true>(cx, prm, w, y, t, dt, e, flag);
  }
  for (int c = 0; c < D; c++) y_io[c] = y[c];
  return result(cx, flag);
}
// evolveHam over a grid (h and the FSAL derivative carried), as hb_traj_evolve<ADAPTIVE> does; out = s rows of D
extern "C" int evolve_rkf45(const double* prm, const double* y0, const double* ts, int s, double* out) {
  double w[S::M];
  S::inertia(prm, w);
  HbCtx cx = ctx();
  int flag = 0;
  double y[D];
  for (int c = 0; c < D; c++) y[c] = out[c] = y0[c];
  HbEvolve<D> e;
  e.h = (ts[1] - ts[0]) / 100;
  e.primed = false;
  double t = ts[0];
  for (int k = 1; k < s; k++) {
    hb_rkf45_to<S, true>(cx, prm, w, y, t, ts[k], e, flag);
    for (int c = 0; c < D; c++) out[k * D + c] = y[c];
  }
  return result(cx, flag);
}
extern "C" int config_maps(const double* prm, const double* q, const double* v, double* p_out, double* v_back, double* U) {
  double w[S::M];
  S::inertia(prm, w);
  HbCtx cx = ctx();
  int flag = 0;
  hb_momenta<S, true>(cx, prm, w, q, v, p_out);
  hb_velocities<S, true, true>(cx, prm, w, q, p_out, v_back, *U, flag);
  return result(cx, flag);
}
// primitives
extern "C" int sincos_fast(double x, double* s, double* c) { HbCtx cx = ctx(); hb_sincos<true>(cx, x, s, c); return cx.oob ? 1 : 0; }
extern "C" double rcp_fast(double d) { return hb_rcp(d); }
extern "C" int exp_fast(double x, double* e) { HbCtx cx = ctx(); *e = hb_exp<true>(cx, x); return cx.oob ? 1 : 0; }
template <int NN> static int spd(const double* A_packed, const double* b, double* x) {
  double A[NN * (NN + 1) / 2];
  for (int i = 0; i < NN * (NN + 1) / 2; i++) A[i] = A_packed[i];
  int minpiv = 0x7fffffff;
  hb_spd_solve<NN, true>(A, b, x, minpiv);
  return minpiv < 0x00100000 ? HB_FLAG_NOT_SPD : 0;
}
extern "C" int spd_solve(int n, const double* A_packed, const double* b, double* x) {
  switch (n) {
    case 1: return spd<1>(A_packed, b, x);
    case 2: return spd<2>(A_packed, b, x);
    case 3: return spd<3>(A_packed, b, x);
    case 4: return spd<4>(A_packed, b, x);
    case 5: return spd<5>(A_packed, b, x);
    case 7: return spd<7>(A_packed, b, x);
    case 12: return spd<12>(A_packed, b, x);
    default: return -1;
  }
}

// ---- whole kernels, one emulated thread after another -----------------------------------------------------------------
// HB_DEFINE_KERNELS gives the nine per-system kernels as ordinary functions hk_<kind>(HbKArgs); run_kernel() walks the
// (blockIdx, threadIdx) space sequentially — each call runs that thread's whole grid-stride loop, including the prefetch
// of its next Phase, the out-of-line slow retry and the flag / finite bookkeeping.
HB_DEFINE_KERNELS(SYS_NAME, hk)
static void hk_init_random(const HbKArgs a) { hb_body_init_random(a); }

extern "C" int run_kernel(int kid, const double* in, double* out, int* flags, const double* ts, long long n_traj, double dt, int nsteps,
                          int layout, int s, int substeps, const double* prm, int grid, int block) {
  HbKArgs a;
  std::memset(&a, 0, sizeof a);
  a.in = in; a.out = out; a.flags = flags; a.ts = ts; a.N = n_traj; a.dt = dt; a.dt6 = dt / 6.0; a.dth = 0.5 * dt;
  a.nsteps = nsteps; a.layout = layout; a.s = s; a.substeps = substeps;
  for (int k = 0; k < HB_MAXP; k++) a.prm[k] = prm[k];
  blockDim.x = (unsigned)block; gridDim.x = (unsigned)grid;
  for (unsigned b = 0; b < (unsigned)grid; b++)
    for (unsigned t = 0; t < (unsigned)block; t++) {
      blockIdx.x = b; threadIdx.x = t;
      switch (kid) {
        case HB_K_STEP_RK4: hk_step_rk4(a); break;
        case HB_K_STEP_RKF45: hk_step_rkf45(a); break;
        case HB_K_EVOLVE_RK4: hk_evolve_rk4(a); break;
        case HB_K_EVOLVE_RKF45: hk_evolve_rkf45(a); break;
        case HB_K_HAM_EQS: hk_ham_eqs(a); break;
        case HB_K_TO_PHASE: hk_to_phase(a); break;
        case HB_K_FROM_PHASE: hk_from_phase(a); break;
        case HB_K_ENERGIES: hk_energies(a); break;
        case HB_K_UPOS: hk_upos(a); break;
        default: return -1;
      }
    }
  blockIdx.x = threadIdx.x = 0; blockDim.x = 128; gridDim.x = 1;
  return 0;
}
extern "C" void run_init_random(double* out, long long n_traj, int d, int layout, unsigned long long seed, long long first,
                                const double* lo, const double* hi, int grid, int block) {
  HbKArgs a;
  std::memset(&a, 0, sizeof a);
  a.out = out; a.N = n_traj; a.nsteps = d; a.layout = layout; a.seed = seed; a.first = first;
  for (int c = 0; c < d; c++) { a.prm[c] = lo[c]; a.prm[d + c] = hi[c]; }
  blockDim.x = (unsigned)block; gridDim.x = (unsigned)grid;
  for (unsigned b = 0; b < (unsigned)grid; b++)
    for (unsigned t = 0; t < (unsigned)block; t++) { blockIdx.x = b; threadIdx.x = t; hk_init_random(a); }
  blockIdx.x = threadIdx.x = 0; blockDim.x = 128; gridDim.x = 1;
}
