/*
 * hamilton_b200.h — C ABI of the B200-native batched Hamiltonian-dynamics engine.
 *
 * This is the drop-in boundary for ONE hot path of mstksg/hamilton's Numeric.Hamilton:
 * the equations of motion (hamEqs) and their time stepping (stepHam / evolveHam), plus the
 * thin Config<->Phase maps and energy observables every caller of that path needs.
 * The reference has no FFI of its own (pure Haskell on ad + hmatrix + hmatrix-gsl); each entry
 * point below names the Haskell export it replaces (file:line relative to the reference tree).
 * A Haskell shim binds these with `foreign import ccall`.
 *
 * Conventions
 *   n = generalized coordinates, m = underlying Cartesian coordinates (System m n,
 *   src/Numeric/Hamilton.hs:160-169).  All numbers are IEEE fp64.
 *   A Phase is 2n doubles [q_0..q_{n-1}, p_0..p_{n-1}] — exactly the packing evolveHam hands to
 *   GSL (src/Numeric/Hamilton.hs:457-458).  A Config is [q.., v..] the same way (:103-113).
 *   Batches hold N independent trajectories in one of two layouts:
 *     HB_LAYOUT_AOS  y[i*2n + c]   (array of Phases; what a Storable vector of Phase looks like)
 *     HB_LAYOUT_SOA  y[c*N  + i]   (component-major; one coalesced stream per component)
 *   Buffers live either in host memory (HB_MEM_HOST, blocking: page-locked mapped buffers are read and
 *   written in place by the kernel over PCIe, anything else is staged H2D/D2H by the library) or in
 *   device memory of the current CUDA device (HB_MEM_DEVICE: asynchronous on `stream`).
 *   Every function returns hb_status (0 = HB_OK); hb_last_error() gives a thread-local message.
 *   Numerical failures never abort: they set per-trajectory bits in the optional `flags` array
 *   (HB_FLAG_*), the analogue of hmatrix's `inv` exception / GSL error in the reference.
 *   There is NO CPU fallback: without a CUDA device every compute entry point returns
 *   HB_ERR_NO_DEVICE.
 */
#ifndef HAMILTON_B200_H
#define HAMILTON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 2
#define HB_MAX_N 16      /* generalized coordinates per system            */
#define HB_MAX_M 48      /* Cartesian coordinates per system              */
#define HB_MAX_PARAMS 32 /* runtime parameters (HB_OP_PARAM leaves)       */

typedef int32_t hb_status;
enum {
  HB_OK = 0,
  HB_ERR_INVALID = 1,    /* bad argument (null pointer, size, unknown id)                         */
  HB_ERR_NO_DEVICE = 2,  /* no usable CUDA device / driver: there is no CPU fallback              */
  HB_ERR_CUDA = 3,       /* CUDA runtime error (message in hb_last_error)                         */
  HB_ERR_COMPILE = 4,    /* tape -> CUDA source -> NVRTC failed (log in hb_last_error)            */
  HB_ERR_TAPE = 5,       /* malformed tape (forward reference, bad opcode, bad arity)             */
  HB_ERR_NUMERIC = 6,    /* single-trajectory call hit a numerical failure (flags != 0); mirrors
                            the reference's `error`/exception behaviour for one Phase             */
  HB_ERR_UNSUPPORTED = 7
};

/* Per-trajectory failure bits (OR-ed into flags[i]). */
enum {
  HB_FLAG_NOT_SPD = 1,      /* JtWJ pivot <= 0: reference `inv jmj` (src/Numeric/Hamilton.hs:381) would throw/blow up */
  HB_FLAG_NONFINITE = 2,    /* state became NaN/Inf                                                  */
  HB_FLAG_STEP_FAILED = 4   /* RKF45 controller could not shrink h further (GSL_FAILURE in evolve_apply) */
};

typedef enum { HB_LAYOUT_AOS = 0, HB_LAYOUT_SOA = 1 } hb_layout;
typedef enum { HB_MEM_HOST = 0, HB_MEM_DEVICE = 1 } hb_memspace;

/* Integrators.  HB_INTEG_RK4 is the classical fixed-step RK4 BASELINE.json's metric counts.
 * HB_INTEG_RKF45_GSL reproduces what the reference's stepHam/evolveHam really run:
 * GSL's adaptive rkf45 + standard controller via hmatrix-gsl `odeSolveV RKf45 hi eps eps`
 * (src/Numeric/Hamilton.hs:445-448). */
typedef enum { HB_INTEG_RK4 = 0, HB_INTEG_RKF45_GSL = 1 } hb_integrator;

/* ---------------------------------------------------------------------------------------------
 * Tapes: how a `forall a. RealFloat a => Vector n a -> Vector m a` (mkSystem's rank-2 argument,
 * src/Numeric/Hamilton.hs:212-215) crosses a C boundary.  The host language instantiates `a` at a
 * tracing number type and records a Wengert list; node k may only refer to nodes < k.
 * ------------------------------------------------------------------------------------------- */
typedef enum {
  HB_OP_INPUT = 0,   /* a = input index                                        */
  HB_OP_CONST = 1,   /* c = literal                                            */
  HB_OP_PARAM = 2,   /* a = index into the system's runtime parameter vector   */
  HB_OP_ADD = 3, HB_OP_SUB = 4, HB_OP_MUL = 5, HB_OP_DIV = 6,        /* (a, b) */
  HB_OP_NEG = 7, HB_OP_RECIP = 8, HB_OP_ABS = 9, HB_OP_SIGNUM = 10,  /* (a)    */
  HB_OP_SQRT = 11, HB_OP_EXP = 12, HB_OP_LOG = 13,
  HB_OP_SIN = 14, HB_OP_COS = 15, HB_OP_TAN = 16,
  HB_OP_ASIN = 17, HB_OP_ACOS = 18, HB_OP_ATAN = 19,
  HB_OP_SINH = 20, HB_OP_COSH = 21, HB_OP_TANH = 22,
  HB_OP_ASINH = 23, HB_OP_ACOSH = 24, HB_OP_ATANH = 25,
  HB_OP_POW = 26,    /* a ** b  (Floating (**))                                 */
  HB_OP_POWI = 27,   /* a ^ k, integer k stored in c  (Num (^), Fractional (^^)) */
  HB_OP_ATAN2 = 28,  /* atan2 a b (RealFloat)                                   */
  HB_OP__COUNT = 29
} hb_opcode;

typedef struct hb_op {
  int32_t op;   /* hb_opcode */
  int32_t a;    /* first operand node / input index / param index */
  int32_t b;    /* second operand node (binary ops), else 0       */
  int32_t _pad;
  double c;     /* literal (CONST) or integer exponent (POWI)     */
} hb_op;

typedef struct hb_tape {
  int32_t n_in;          /* number of inputs                          */
  int32_t n_ops;
  const hb_op* ops;
  int32_t n_out;
  const int32_t* outs;   /* node index of each output                 */
} hb_tape;

typedef struct hb_system hb_system; /* opaque, immutable after creation, shareable across threads;
                                       replaces `System m n` (src/Numeric/Hamilton.hs:160-169) */

/* Built-in systems = the reference's example fixtures (app/Examples.hs:61-183) plus the two
 * BASELINE.json benchmark systems that extend them.  They are compiled ahead of time into the
 * library (no NVRTC needed).  `params` may be NULL for the reference's CLI defaults. */
typedef enum {
  HB_SYS_PENDULUM = 0,        /* System 2 1, app/Examples.hs:61-73;   params: none                     */
  HB_SYS_DOUBLE_PENDULUM = 1, /* System 4 2, app/Examples.hs:75-94;   params: m1, m2   (defaults 1, 1) */
  HB_SYS_ROOM = 2,            /* System 2 2, app/Examples.hs:96-116;  params: none                     */
  HB_SYS_TWO_BODY = 3,        /* System 4 2, app/Examples.hs:118-142; params: m1, m2   (5, 0.5)        */
  HB_SYS_SPRING = 4,          /* System 3 3, app/Examples.hs:144-162; params: mB, mW, k (2, 1, 10)     */
  HB_SYS_BEZIER = 5,          /* System 2 1, app/Examples.hs:164-183; params: 5 control points x,y (defaults :350) */
  HB_SYS_TRIPLE_PENDULUM = 6, /* System 6 3; params: m1,m2,m3,l1,l2,l3 (1,1,1,1,.5,.5) */
  HB_SYS_CHAIN12 = 7,         /* System 24 12; params: none (l=1, m=1, g=5)  */
  HB_SYS_SPRING1D = 8,        /* System 2 1, synthetic; params: k, alpha (10, 0.3) */
  HB_SYS__COUNT = 9
} hb_builtin;

/* ---- library / device -------------------------------------------------------------------- */
int32_t hb_abi_version(void);
const char* hb_last_error(void);              /* thread-local, never NULL                         */
hb_status hb_device_count(int32_t* count);    /* HB_ERR_NO_DEVICE when CUDA is unusable            */
hb_status hb_set_device(int32_t device);      /* selects the CUDA device for this thread           */

/* ---- system construction (replaces mkSystem / mkSystem') ---------------------------------- */
hb_status hb_system_builtin(hb_builtin id, const double* params, int32_t n_params, hb_system** out);

/* mkSystem (src/Numeric/Hamilton.hs:201-225): f: R^n -> R^m, u: R^n -> R.
 * mkSystem' (:238-254): u_on_cartesian != 0 and u: R^m -> R; the library forms u . f.
 * The tapes are differentiated symbolically (forward mode, second order — what `jacobianT`,
 * `hessianF` and `grad` do at run time in the reference, :221-224), specialised to CUDA source
 * and compiled with NVRTC for the current device's architecture. */
hb_status hb_system_from_tape(int32_t m, int32_t n, const double* inertia /* m */,
                              const hb_tape* f, const hb_tape* u, int32_t u_on_cartesian,
                              const double* params, int32_t n_params, hb_system** out);
void hb_system_free(hb_system* sys);
hb_status hb_system_dims(const hb_system* sys, int32_t* m, int32_t* n);
/* Generated CUDA source of the system's derivative code (diagnostics / docs).  Returns bytes
 * needed (including NUL); copies at most `cap`. */
size_t hb_system_source(const hb_system* sys, char* buf, size_t cap);
/* Runtime parameter vector the generated code reads as prm[k] (for built-ins: derived from the user
 * parameters, e.g. two-body stores m1, m2, -(m2/mT), m1/mT, m1*m2).  Returns the count; copies at most `cap`. */
int32_t hb_system_params(const hb_system* sys, double* buf, int32_t cap);

/* ---- batched hot path ---------------------------------------------------------------------
 * All arrays hold N trajectories in `layout`, in `mem`; `stream` is a cudaStream_t passed as
 * void* (NULL = default stream) and only matters for HB_MEM_DEVICE, where calls are asynchronous.
 * `flags` (N x int32, same memspace, may be NULL) receives HB_FLAG_* bits; the caller zeroes it. */

/* hamEqs (src/Numeric/Hamilton.hs:370-387): dy[i] = (dq/dt, dp/dt) at y[i]. */
hb_status hb_batch_ham_eqs(const hb_system* sys, int64_t N, hb_layout layout, hb_memspace mem,
                           const double* y, double* dy, int32_t* flags, void* stream);

/* stepHam iterated (src/Numeric/Hamilton.hs:390-402): advance every trajectory by `nsteps`
 * steps of size dt with `integ`.  RK4: nsteps classical RK4 updates.  RKF45_GSL: nsteps
 * independent `stepHam dt` calls, i.e. a fresh adaptive solve over (0, dt) with h0 = dt/100,
 * eps = 1.49012e-08 each (:445-448).  y_in may equal y_out. */
hb_status hb_batch_step(const hb_system* sys, hb_integrator integ, double dt, int32_t nsteps,
                        int64_t N, hb_layout layout, hb_memspace mem,
                        const double* y_in, double* y_out, int32_t* flags, void* stream);

/* evolveHam (src/Numeric/Hamilton.hs:433-462) for N trajectories sharing one time grid ts[0..s):
 * out[k] is the batch at ts[k] (out is s consecutive batches in `layout`); out[0] = y0, as the
 * reference's first row is the initial state.  RKF45_GSL carries h and the FSAL derivative across
 * grid points exactly like hmatrix-gsl's loop; RK4 takes `rk4_substeps` equal steps per interval. */
hb_status hb_batch_evolve(const hb_system* sys, hb_integrator integ, int32_t rk4_substeps,
                          int64_t N, hb_layout layout, hb_memspace mem,
                          const double* y0, const double* ts, int32_t s, double* out,
                          int32_t* flags, void* stream);

/* toPhase / momenta (:262-284): c = [q, v] -> y = [q, p].   fromPhase / velocities (:316-337). */
hb_status hb_batch_to_phase(const hb_system* sys, int64_t N, hb_layout layout, hb_memspace mem,
                            const double* c, double* y, void* stream);
hb_status hb_batch_from_phase(const hb_system* sys, int64_t N, hb_layout layout, hb_memspace mem,
                              const double* y, double* c, int32_t* flags, void* stream);

/* Observables on Phases: out4[i*4 + {0,1,2,3}] = keP, pe, hamiltonian (T+U), lagrangian (T-U)
 * (:341-361, :182-186, :301-309).  `out4` is always AOS N x 4 in `mem`. */
hb_status hb_batch_energies(const hb_system* sys, int64_t N, hb_layout layout, hb_memspace mem,
                            const double* y, double* out4, int32_t* flags, void* stream);

/* underlyingPos (:174-178): x[i] = f(q[i]); q is N x n, x is N x m, both in `layout`. */
hb_status hb_batch_underlying_pos(const hb_system* sys, int64_t N, hb_layout layout,
                                  hb_memspace mem, const double* q, double* x, void* stream);

/* Counter-based synthetic initial Phases: component c of trajectory i is
 * lo[c] + (hi[c]-lo[c]) * u,  u = (splitmix64(seed + 2n*(first+i) + c) >> 11) * 2^-53.
 * Device-side generation so multi-GPU shards need no scatter. y in HB_MEM_DEVICE only. */
hb_status hb_batch_init_random(const hb_system* sys, uint64_t seed, int64_t first, int64_t N,
                               hb_layout layout, const double* lo, const double* hi /* 2n each, host */,
                               double* y_device, void* stream);

/* ---- multi-GPU ensembles ----------------------------------------------------
 * Trajectories are independent (stepHam is a pure function of each Phase, src/Numeric/Hamilton.hs:390-399), so an
 * ensemble of N initial conditions is block-split over the GPUs of THIS process: device g of `ndev` owns trajectories
 * [g*N/ndev, (g+1)*N/ndev) in its own memory (array of Phases), steps them with no communication, and ONE NCCL
 * all-gather over NVLink assembles the final Phases on every device.  Single process, one host thread per device inside
 * every call (a Haskell host links with -threaded, hamilton.cabal:54; no MPI / torchrun needed).
 * All calls are blocking unless stated; the ensemble may be used from any one thread at a time. */
typedef struct hb_ensemble hb_ensemble;

/* devices: ndev CUDA device ordinals, or NULL for 0..ndev-1.  Allocates the shards (two N/ndev x 2n buffers per device
 * that swap roles every launch), one stream per device and the NCCL communicators (ncclCommInitAll).
 * HB_ERR_UNSUPPORTED when libnccl cannot be loaded and ndev > 1. */
hb_status hb_ensemble_create(const hb_system* sys, int32_t ndev, const int32_t* devices, int64_t N, hb_ensemble** out);
void hb_ensemble_free(hb_ensemble* ens);
hb_status hb_ensemble_dims(const hb_ensemble* ens, int32_t* ndev, int64_t* N, int64_t* first /* ndev+1 shard offsets, may be NULL */);

/* Initial Phases: generated on every device from the counter-based RNG of hb_batch_init_random with GLOBAL trajectory
 * indices (no scatter), or uploaded from one host array of N Phases (array of Phases). */
hb_status hb_ensemble_init_random(hb_ensemble* ens, uint64_t seed, const double* lo, const double* hi /* 2n each */);
hb_status hb_ensemble_upload(hb_ensemble* ens, const double* y_host /* N x 2n */);

/* `launches` x hb_batch_step(integ, dt, nsteps) on every shard concurrently (state read from and written to HBM every
 * launch).  Returns when every device has finished; *gpu_ms (may be NULL) = device time, max over devices. */
hb_status hb_ensemble_step(hb_ensemble* ens, hb_integrator integ, double dt, int32_t nsteps, int32_t launches, double* gpu_ms);

/* One ncclAllGather of the current Phases: afterwards EVERY device holds all N Phases (hb_ensemble_gathered gives the
 * device pointers); y_host (may be NULL) additionally receives them from device 0.  *gpu_ms = device time of the
 * collective, max over devices.  The first call allocates and registers (ncclCommRegister) the N x 2n receive buffers. */
hb_status hb_ensemble_gather(hb_ensemble* ens, double* y_host /* N x 2n or NULL */, double* gpu_ms);
/* Device pointers: the shard of device index g (Ng x 2n, current state) / its gathered copy (N x 2n, valid after gather). */
hb_status hb_ensemble_shard(const hb_ensemble* ens, int32_t g, double** y_device, int64_t* n_shard);
hb_status hb_ensemble_gathered(const hb_ensemble* ens, int32_t g, double** y_device);
/* Per-trajectory HB_FLAG_* bits accumulated since creation, all N, to host. */
hb_status hb_ensemble_flags(hb_ensemble* ens, int32_t* flags_host /* N */);

/* ---- single-trajectory mirrors of the Haskell API (host pointers, run on the GPU with N = 1).
 * A numerical failure returns HB_ERR_NUMERIC, the analogue of the reference's `error`. -------- */
hb_status hb_underlying_pos(const hb_system* sys, const double* q, double* x);      /* :174-178 */
hb_status hb_pe(const hb_system* sys, const double* q, double* u);                  /* :182-186 */
hb_status hb_momenta(const hb_system* sys, const double* q, const double* v, double* p); /* :262-269 */
hb_status hb_velocities(const hb_system* sys, const double* q, const double* p, double* v); /* :316-324 */
hb_status hb_ke_c(const hb_system* sys, const double* q, const double* v, double* t);  /* keC :288-296 */
hb_status hb_ke_p(const hb_system* sys, const double* q, const double* p, double* t);  /* keP :341-349 */
hb_status hb_lagrangian(const hb_system* sys, const double* q, const double* v, double* l);  /* :301-309 */
hb_status hb_hamiltonian(const hb_system* sys, const double* q, const double* p, double* h); /* :353-361 */
hb_status hb_ham_eqs(const hb_system* sys, const double* q, const double* p,
                     double* dq, double* dp);                                       /* :370-387 */
hb_status hb_step_ham(const hb_system* sys, double r, const double* q, const double* p,
                      double* q_out, double* p_out);                                /* :390-402 */
hb_status hb_evolve_ham(const hb_system* sys, const double* q0, const double* p0,
                        const double* ts, int32_t s, double* out /* s x 2n */);     /* :433-462 */
hb_status hb_step_ham_c(const hb_system* sys, double r, const double* q, const double* v,
                        double* q_out, double* v_out);                              /* stepHamC :505-515 */
hb_status hb_evolve_ham_c(const hb_system* sys, const double* q0, const double* v0,
                          const double* ts, int32_t s, double* out /* s x 2n */);   /* evolveHamC :488-498 */

#ifdef __cplusplus
}
#endif
#endif /* HAMILTON_B200_H */
