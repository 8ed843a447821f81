// hb_engine.cuh — hand-written sm_100a device code of the batched Hamiltonian engine.
//
// One thread owns one phase-space trajectory and keeps its whole state (q, p, the Jacobian
// non-zeros, the packed mass matrix and its LDL^T factor, the RK stage vectors) in registers;
// the FP64 pipe, not HBM, is the binding resource for this path (DESIGN.md "Roofline"), so the
// design goal is the minimum number of DFMA-class instructions per RHS evaluation and 100 %
// useful lanes, with coalesced 16-byte loads/stores of the Phase arrays on either side.
//
// The engine is generic over a `Sys` description that the host-side system compiler
// (csrc/sysgen.cpp) emits from the user's tapes — what mkSystem's `jacobianT`/`hessianF`/`grad`
// closures are in the reference (src/Numeric/Hamilton.hs:217-225), resolved at System-construction
// time instead of on every RHS call:
//
//   struct Sys {
//     static constexpr int M, N;        // System m n
//     static constexpr int NJ, NH;      // structural non-zeros of J (m x n) and of the Hessian tensor
//     static constexpr int jidx(i, j);  // position of J[i][j] in the packed list, -1 if J[i][j] == 0
//     static constexpr int jrow(e), jcol(e);
//     static constexpr int hrow(e), hj(e), hk(e);   // H entry e is d2 f_hrow / dq_hj dq_hk, hj <= hk
//     static constexpr int NG, hgrp(e), gj(g), gk(g);  // H entries grouped by their (j, k) pair
//     static constexpr bool TRIG;       // uses sin/cos (needs the shared-memory table)
//     static void inertia(prm, w[M]);
//     template <bool FAST> static void derivs(cx, prm, q, Jv[NJ], Hv[NH], gU[N]);   // everything hamEqs needs
//     template <bool FAST> static void jac(cx, prm, q, Jv[NJ]);
//     template <bool FAST> static void jac_pot(cx, prm, q, Jv[NJ], U);
//     template <bool FAST> static void pos(cx, prm, q, x[M]);
//   };
// FAST selects the hand-written fp64 primitives below; FAST=false is the out-of-line retry with
// libdevice math for the rare trajectory whose arguments leave the fast primitives' domain.
//
// This header is compiled ahead of time by nvcc for the built-in systems and at run time by
// NVRTC for tape systems, so it must not include any standard header.
#pragma once

#define HB_MAXP 64
// Threads per CTA, fixed per system size (the host launches exactly this): 128 for small systems whose whole
// state lives in registers and for n >= 8, where the RK vectors are staged in dynamic shared memory (HB_DYN_DOUBLES).
// Small systems: the host picks the CTA size per launch (csrc/runtime.cpp pick_block): one CTA per SM of 8..16 warps when the
// kernel's registers allow it — one 32 KB table image staged per SM per launch — with the warp count chosen so that the
// batch's tiles divide into (nearly) whole rounds.  HB_BLOCK_SMALL is the upper bound the kernels are compiled for
// (__launch_bounds__: 512 threads leave every kernel 128 registers).
#ifndef HB_BLOCK_SMALL
#define HB_BLOCK_SMALL 512
#endif
#define HB_BLOCK_OF(NCOORD) ((NCOORD) >= HB_BIG_N ? 128 : HB_BLOCK_SMALL)
// Large systems (n >= HB_BIG_N) keep the RK4 vectors (y, acc, stage input: 3 * 2n doubles per thread) and, for
// symbolically compiled systems, the values hpre hands to hpost (NE doubles per thread) in DYNAMIC shared memory,
// column per thread (conflict-free), so that the registers are left to the mass matrix and its factor.
#define HB_BIG_N 8
#define HB_DYN_DOUBLES(NCOORD, NE_) ((NCOORD) >= HB_BIG_N ? 3 * 2 * (NCOORD) + (NE_) : 0)
// Dynamic shared memory of every kernel (hb_dsm), in this order — csrc/runtime.cpp dyn_smem_bytes() computes the same sizes:
//   small systems (n < HB_BIG_N):  [ table image (sin/cos + 2^(j/64)): HB_TAB_BYTES, systems with sin/cos/exp only (Sys::TRIG) ]  [ stage: 2 x DIN doubles per
//                                  thread, the cp.async landing zone of the thread's next Phase: stepping kernels of HEAVY
//                                  systems only ]  [ xp ]
//   large systems:                 [ RK vectors + parked values: HB_DYN_DOUBLES per thread ]  [ xp ]
//   xp = DOUT doubles per thread, the warps' transpose buffers for host-memory stores; present only when a.layout == 2
//   and DOUT is even and <= HB_WSTORE_MAXD.
#define HB_TAB_BYTES 33408   // sizeof(HbTab) rounded up to 128 (static_assert below)
// Upper bound of the CTA size a kernel may be launched with (= __launch_bounds__).  A sweep of CTA sizes 128..384 x
// resident waves 1..4 on B200 (profiles/r1d/sweep.txt) was flat within run-to-run noise, so it is the default size.
#define HB_MAXBLOCK_OF(NCOORD) HB_BLOCK_OF(NCOORD)
#define HB_BLOCK 128      // system-independent kernels (init_random)
#ifndef HB_WSTORE_MAXD
#define HB_WSTORE_MAXD 8   // widest record (doubles) the warp-transposed host store (layout 2) handles: 4 KB of shared memory per CTA
#endif
#define HB_DEV __device__ __forceinline__

#ifndef HAMILTON_B200_H   // same values as the enum in include/hamilton_b200.h (not includable under NVRTC)
#define HB_FLAG_NOT_SPD 1
#define HB_FLAG_NONFINITE 2
#define HB_FLAG_STEP_FAILED 4
#endif

// Uniform kernel-argument block (passed by value: lands in the constant bank, so `prm[k]`
// and dt become direct constant operands of DFMA).
struct HbKArgs {
  const double* in;       // input batch
  double* out;            // output batch
  int* flags;             // optional per-trajectory HB_FLAG_* (may be null)
  const double* ts;       // evolve: time grid on the device
  long long N;            // trajectories
  double dt;              // step size
  double dt6;             // dt / 6 (host-computed: keeps a full-precision division out of every thread)
  double dth;             // dt / 2 (host-computed: a kernel-argument operand instead of a register computed per thread)
  int nsteps;             // steps per launch
  int layout;             // 0 = AOS y[i*D+c], 1 = SOA y[c*N+i]
  int s;                  // evolve: number of grid points
  int substeps;           // evolve/RK4: equal sub-steps per grid interval
  unsigned long long seed;
  long long first;
  int host_io;            // 1: `in` / `out` are page-locked HOST memory accessed over PCIe (no L2 prefetches of them)
  int contiguous;         // 1: CTA b starts at tiles b W .. b W + W - 1 (multi-wave grids of long kernels), 0: tiles spread over CTAs
  double prm[HB_MAXP];    // runtime parameters (HB_OP_PARAM leaves); for init_random: lo[0..D), hi[0..D)
};

// ---------------------------------------------------------------------------- batch I/O ----
// I = index type: `unsigned` in the kernels (a launch holds < 2^31 trajectories, so i * D * 8 is ONE IMAD.WIDE.U32 instead
// of a 64-bit multiply-add chain), `long long` in the out-of-line slow path.
template <int D, class I>
HB_DEV void hb_load(const double* __restrict__ base, I i, I N, int layout, double (&y)[D]) {
  if (layout != 1) {
    if constexpr (D % 2 == 0) {   // 16-byte vector loads: a Phase is 2n doubles, always even
      const double2* p = reinterpret_cast<const double2*>(base + (size_t)i * D);
#pragma unroll
      for (int c = 0; c < D / 2; c++) { double2 v = p[c]; y[2 * c] = v.x; y[2 * c + 1] = v.y; }
    } else {
#pragma unroll
      for (int c = 0; c < D; c++) y[c] = base[(size_t)i * D + c];
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; c++) y[c] = base[(size_t)c * N + i];   // one coalesced stream per component
  }
}
// layout 2 (internal; set by the host when `out` is page-locked HOST memory written in place over PCIe): array of Phases
// like layout 0, but a warp transposes its 32 Phases through shared memory so that every store instruction writes 512
// contiguous bytes.  A thread-per-Phase store (16 bytes at a 2n*8-byte stride) reaches the host as partial-sector
// writes: 12.7 GB/s measured against 53.6 GB/s for the transposed form (profiles/r1h/zc.txt).
template <int D, class I>
HB_DEV void hb_store(double* __restrict__ base, I i, I N, int layout, const double (&y)[D], double* xp = nullptr) {
  if constexpr (D % 2 == 0 && D <= HB_WSTORE_MAXD) {
    if (layout == 2) {
      // xp: this warp's transpose buffer (32 * D doubles of the kernel's dynamic shared memory; null on the slow path)
      if (xp != nullptr && __activemask() == 0xffffffffu) {   // whole warp here together: lanes hold 32 consecutive trajectories
        const int lane = threadIdx.x & 31;
        double2* w = reinterpret_cast<double2*>(xp);
#pragma unroll
        for (int c = 0; c < D / 2; c++) w[lane * (D / 2) + c] = make_double2(y[2 * c], y[2 * c + 1]);
        __syncwarp();
        double2* g = reinterpret_cast<double2*>(base + (size_t)(i - lane) * D);
#pragma unroll
        for (int k = 0; k < D / 2; k++) g[32 * k + lane] = w[32 * k + lane];
        __syncwarp();
        return;
      }
      layout = 0;   // ragged warp (batch tail, or a lane left for the slow path): plain per-thread stores
    }
  }
  if (layout != 1) {
    if constexpr (D % 2 == 0) {
      double2* p = reinterpret_cast<double2*>(base + (size_t)i * D);
#pragma unroll
      for (int c = 0; c < D / 2; c++) p[c] = make_double2(y[2 * c], y[2 * c + 1]);
    } else {
#pragma unroll
      for (int c = 0; c < D; c++) base[(size_t)i * D + c] = y[c];
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; c++) base[(size_t)c * N + i] = y[c];
  }
}
// L2 prefetch of one trajectory's input record (every 32-byte sector of it).  A prefetch into L2 — the memory-side cache,
// the point of coherence — can never make stale data visible, so the kernels issue it for their first two rounds BEFORE
// griddepcontrol.wait: the DRAM latency of the first loads hides under the previous kernel's tail.
#ifndef HB_PRE_L2
#define HB_PRE_L2 1
#endif

template <int D, class I>
HB_DEV void hb_prefetch_l2(const double* base, I i, I N, int layout) {
#ifdef HB_HOST_EMU
  (void)base; (void)i; (void)N; (void)layout;
#else
  if (layout != 1) {
#pragma unroll
    for (int c = 0; c < D; c += 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)i * D + c));
  } else {
#pragma unroll
    for (int c = 0; c < D; c++) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)c * N + i));
  }
#endif
}
// Shared-window address of a shared-memory object, computed ONCE: the volatile asm keeps ptxas from rematerialising the
// generic->shared conversion (S2R SR_CgaCtaId + MOV + LEA) in front of every use inside the trajectory loop.
HB_DEV unsigned hb_smem_addr(const void* p) {
#ifdef HB_HOST_EMU
  (void)p; return 0u;
#else
  unsigned r;
  asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\tcvt.u32.u64 %0, t;\n\t}" : "=r"(r) : "l"(p));
  return r;
#endif
}
// cp.async (LDGSTS) staging of one array-of-records trajectory: D/2 16-byte copies straight into shared memory, no
// destination registers — so the copy of the NEXT Phase can be issued a whole trajectory ahead.  (A register prefetch is
// sunk by ptxas to the end of the current step to shorten its live range, and the DRAM latency it should hide comes back
// as a long-scoreboard stall on the first use: 35 % of all warp samples, profiles/r2a.)  Slot layout: double2 c of thread t
// at slot[(c * blockDim + t)]: conflict-free LDS.128.
// slot: this thread's first double2 of the stage buffer (generic pointer, used by the host emulation);  slot_s: its
// shared-window address;  cstride: bytes between the thread's consecutive double2 (= 16 * blockDim)
template <int D>
HB_DEV void hb_async_load(double* slot, unsigned slot_s, unsigned cstride, const double* src) {
#ifdef HB_HOST_EMU
  (void)slot_s;
#pragma unroll
  for (int c = 0; c < D / 2; c++) { slot[c * (cstride / 8)] = src[2 * c]; slot[c * (cstride / 8) + 1] = src[2 * c + 1]; }
#else
  (void)slot;
#pragma unroll
  for (int c = 0; c < D / 2; c++)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot_s + c * cstride), "l"(src + 2 * c) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int D>
HB_DEV void hb_async_read(const double* slot, unsigned slot_s, unsigned cstride, double (&y)[D]) {
#ifdef HB_HOST_EMU
  (void)slot_s;
#pragma unroll
  for (int c = 0; c < D / 2; c++) { y[2 * c] = slot[c * (cstride / 8)]; y[2 * c + 1] = slot[c * (cstride / 8) + 1]; }
#else
  (void)slot;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int c = 0; c < D / 2; c++)
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(y[2 * c]), "=d"(y[2 * c + 1]) : "r"(slot_s + c * cstride) : "memory");
#endif
}
HB_DEV bool hb_finite(double x) { return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000; }
// HB_HOST_EMU: tests/host_engine_harness.cpp compiles this header as HOST code (one "thread", CUDA built-ins stubbed) so
// that hamEqs, the SPD solves, RK4, the GSL-RKF45 stepper/controller and the fast sincos/reciprocal can be checked
// against the oracle without a GPU.  Only the inline-PTX sites need an alternative; device builds never define it.
#ifdef HB_HOST_EMU
#define HB_PDL_LAUNCH_DEPENDENTS() ((void)0)
#define HB_PDL_WAIT() ((void)0)
#else
#define HB_PDL_LAUNCH_DEPENDENTS() asm volatile("griddepcontrol.launch_dependents;")
#define HB_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#endif
#ifdef HB_HOST_EMU
static double hb_dsm[1 << 17];
HB_DEV double hb_lds(const double* p) { return *p; }
HB_DEV void hb_sts(double* p, double v) { *p = v; }
#else
extern __shared__ __align__(128) double hb_dsm[];   // dynamic shared memory (layout: HB_DYN_DOUBLES above)
HB_DEV double hb_lds(const double* p) {   // opaque to the optimiser: a parked value is re-read, never kept in a register
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
HB_DEV void hb_sts(double* p, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(v) : "memory");
}
#endif
template <int D>
HB_DEV void hb_copy(const double* src, double (&dst)[D]) {
#pragma unroll
  for (int c = 0; c < D; c++) dst[c] = src[c];
}

// ------------------------------------------------------------------------ fp64 primitives --
// On B200 an FP64 instruction keeps a scheduler's issue port for 2 clocks (3 with three distinct register operands) and
// every other instruction costs one more issue clock on top (profiles/r2a/fp64_issue_model.txt), so the engine is bound
// by the TOTAL instruction count of a step, FP64 instructions weighing double.  The two transcendental-class primitives
// every mechanical system leans on are therefore hand-written, and their bookkeeping counted instruction by instruction.
//
// hb_sincos<FAST=true>: x = k * (2 pi / HB_SC_N) + r with k = rint(x * HB_SC_N / 2 pi) obtained from the magic constant
// 1.5 * 2^52 + 2^31 (no F2I/I2F; the 2^31 bias keeps the low word of t = fma(x, HB_SC_N / 2 pi, magic) non-negative, so
// "argument in the domain" is hi(t) == 0x43380000: ONE LOP3 accumulates the domain check);  (sin, cos)(k 2 pi / HB_SC_N)
// come from a 2048-entry table staged in shared memory (one LDS.128, address = LOP3 + IMAD);  |r| <= pi / 2048, so
// sin r = r + r^3 S1' (one minimax term, truncation error 9e-18) and cos r - 1 = z (C1 + z C2); angle-addition
// recombination: 12 FP64 instructions + 4 others, no branch, no quadrant logic (libdevice sincos: ~24 FP64 + ~28 others).
// Reduction: one FMA against fp64(2 pi / HB_SC_N) — the product is exact inside the FMA, the constant's own rounding
// error (3.9e-17 relative) makes the result sin/cos of x (1 + 3.9e-17): a perturbation of 0.36 ulp of the ARGUMENT, below
// the 0.5 ulp the argument already carries from the RK stage update that produced it.  Max abs error
// 2.5e-16 + 3.9e-17 |x| (checked against mpmath / long double on the host, tests/test_cpu_engine_host.py);
// HB_SC_CW2=1 restores the two-term Cody-Waite reduction (error 2.5e-16 over the whole domain, +1 DFMA per sincos).
// Arguments with |x| >= 2^31 * 2 pi / HB_SC_N (6.6e6; or non-finite) only set cx.oob; the caller then redoes that
// trajectory on the out-of-line slow path (FAST=false: libdevice sincos with Payne-Hanek reduction).
struct HbCtx {
  unsigned tab_s;       // shared-window address of the staged sin/cos table (fast path only)
  unsigned oob;         // non-zero when a fast-path primitive saw an argument outside its domain
  int minpiv;           // smallest high word of any mass-matrix pivot seen (hb_bad_pivot test deferred to the end)
  double* xp;           // this warp's transpose buffer for host-memory stores (layout 2), else null
};
HB_DEV void hb_ctx_reset(HbCtx& cx) { cx.oob = 0; cx.minpiv = 0x7fffffff; }

#include "hb_sincos_tab.cuh"   // hb_kSinCosTab[2048] = {sin, cos}(k pi / 1024), correctly rounded (gen_sincos_tab.py)
#ifndef HB_SC_LOG2
#define HB_SC_LOG2 11         // staged entries over the full circle: 11 = the whole table (32 KB), 9 = every 4th entry (8 KB)
#endif
#ifndef HB_SC_CW2
#define HB_SC_CW2 0           // 1: two-term Cody-Waite reduction
#endif
#define HB_SC_N (1 << HB_SC_LOG2)
#define HB_SC_MAGIC 6755401588539392.0   // 1.5 * 2^52 + 2^31
// The constants sit in the constant bank so that they are DIRECT operands of the FP64 instructions (written as literals
// ptxas rebuilds each one in a register pair per trajectory: 18 extra issue slots per RK4 step of the double pendulum).
#if HB_SC_LOG2 == 11
static __device__ __constant__ double hb_kSC[6] = {
    325.94932345220167,        // 0  1024 / pi
    0.0030679615757712823,     // 1  pi / 1024 (fp64)
    1.195944139792337e-19,     // 2  pi / 1024 - fp64(pi / 1024)
    -0.16666664960671299,      // 3  S1' = -1/6 + 0.87 zmax / 120: minimax for sin r = r + r^3 S1' over |r| <= pi/2048
    0.0,
    HB_SC_MAGIC};              // 5  (unused: see hb_sincos)
#elif HB_SC_LOG2 == 9
static __device__ __constant__ double hb_kSC[6] = {
    81.48733086305042,         // 0  256 / pi
    0.01227184630308513,       // 1  pi / 256 (fp64)
    4.783776559169348e-19,     // 2  pi / 256 - fp64(pi / 256)
    -1.0 / 6.0, 1.0 / 120.0,   // 3, 4  sin r = r + r^3 (S1 + z S2), |r| <= pi/512: r^7/5040 < 1e-17 r
    HB_SC_MAGIC};
#else
#error "HB_SC_LOG2 must be 9 or 11"
#endif

// Shared-memory image of the table + the mbarrier its bulk copy completes on.
struct HbTab {
  double2 e[HB_SC_N];   // sin/cos
  double x2[64];        // 2^(j/64) (hb_exp); directly behind e as in the global image, so that one bulk copy brings both
  unsigned long long bar;
};
static_assert(sizeof(HbTab) <= HB_TAB_BYTES, "HB_TAB_BYTES must hold the table image");
#ifndef HB_TAB_BULK
#define HB_TAB_BULK (HB_SC_LOG2 == HB_SC_TAB_LOG2)   // whole table: ONE cp.async.bulk (TMA, UBLKCP) per CTA instead of a load/store loop
#endif
// Stage the table.  Reads only constant data, so it runs BEFORE griddepcontrol.wait, under the previous kernel's tail.
HB_DEV void hb_tab_issue(HbTab* tab) {
#ifndef HB_HOST_EMU
#if HB_TAB_BULK
  if (threadIdx.x == 0) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&tab->bar), dst = (unsigned)__cvta_generic_to_shared(tab->e);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr unsigned bytes = sizeof(tab->e) + sizeof(tab->x2);
    static_assert(sizeof(HbTabImage) == bytes && sizeof(HbTab) == bytes + 16, "shared image = global image (x2 directly behind e)");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(&hb_kTab), "r"(bytes), "r"(bar) : "memory");
  }
#else
#pragma unroll 1
  for (int t = threadIdx.x; t < HB_SC_N; t += blockDim.x) tab->e[t] = hb_kSinCosTab[t << (HB_SC_TAB_LOG2 - HB_SC_LOG2)];
  if (threadIdx.x < 64) tab->x2[threadIdx.x] = hb_kTab.x2[threadIdx.x];
#endif
  __syncthreads();   // the initialised mbarrier (or the staged entries) are visible to every thread of the CTA
#else
  (void)tab;
#endif
}
HB_DEV void hb_tab_wait(HbTab* tab) {
#if !defined(HB_HOST_EMU) && HB_TAB_BULK
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&tab->bar);
  unsigned done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
  } while (!done);
#else
  (void)tab;
#endif
}

template <bool FAST>
HB_DEV void hb_sincos(HbCtx& cx, double x, double* sp, double* cp) {
  if constexpr (!FAST) {
    sincos(x, sp, cp);
  } else {
    const double t = fma(x, hb_kSC[0], HB_SC_MAGIC);   // literal on purpose: from the constant bank (hb_kSC[5]) ptxas loads it with an LDC per trajectory and the step is 1.2 % slower (profiles/r3c/ab_double_pendulum.txt)
    cx.oob |= (unsigned)__double2hiint(t) ^ 0x43380000u;   // 0 <=> -2^31 <= k < 2^31 (inf/nan/huge arguments land elsewhere)
    double2 sc;   // one LDS.128 with a 32-bit shared address (no generic->shared conversion in the loop)
#ifdef HB_HOST_EMU
    sc = hb_kSinCosTab[(__double2loint(t) & (HB_SC_N - 1)) << (HB_SC_TAB_LOG2 - HB_SC_LOG2)];
#else
    unsigned addr;
    asm("{\n\t.reg .u32 k;\n\tand.b32 k, %1, %2;\n\tmad.lo.u32 %0, k, 16, %3;\n\t}" : "=r"(addr) : "r"(__double2loint(t)), "n"(HB_SC_N - 1), "r"(cx.tab_s));
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(sc.x), "=d"(sc.y) : "r"(addr));
#endif
    const double kf = t - HB_SC_MAGIC;
    double r = fma(-kf, hb_kSC[1], x);
#if HB_SC_CW2
    r = fma(-kf, hb_kSC[2], r);
#endif
    const double z = r * r;
#if HB_SC_LOG2 == 11
    const double sr = fma(z * r, hb_kSC[3], r);                           // sin r
    const double cm = z * fma(z, 1.0 / 24.0, -0.5);                       // cos r - 1
#else
    const double sr = fma(z * r, fma(z, hb_kSC[4], hb_kSC[3]), r);
    const double cm = z * fma(z, 1.0 / 24.0, -0.5);                       // z^3/720 < 8e-17
#endif
    *sp = fma(sc.y, sr, fma(sc.x, cm, sc.x));    // sin(a + r) = sin a + (sin a (cos r - 1) + cos a sin r)
    *cp = fma(-sc.x, sr, fma(sc.y, cm, sc.y));   // cos(a + r) = cos a + (cos a (cos r - 1) - sin a sin r)
  }
}
template <bool FAST> HB_DEV double hb_sin(HbCtx& cx, double x) { double s, c; hb_sincos<FAST>(cx, x, &s, &c); return s; }
template <bool FAST> HB_DEV double hb_cos(HbCtx& cx, double x) { double s, c; hb_sincos<FAST>(cx, x, &s, &c); return c; }

// hb_exp<FAST>: the `exp` of user tapes (the logistic walls of the reference's room / spring / bezier examples,
// app/Examples.hs:601-605).  x = (64 m + j) ln2/64 + r, |r| <= ln2/128:  exp x = 2^m * 2^(j/64) * e^r with 2^(j/64) from the
// staged table and e^r - 1 = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120) (truncation r^6/720 < 3.5e-17): 11 FP64 + 5 other
// instructions against ~30 for libdevice's exp (degree-11 polynomial + range handling).  Two-term Cody-Waite reduction with
// k * L1 exact (L1 has 19 trailing zero bits, |k| < 2^16).  Error < 2.5e-16 relative.  Domain |x| < 512 (result normal, no
// overflow/denormal handling); outside (and NaN) only sets cx.oob and the trajectory is redone with libdevice's exp.
template <bool FAST>
HB_DEV double hb_exp(HbCtx& cx, double x) {
  if constexpr (!FAST) {
    return exp(x);
  } else {
    cx.oob |= (unsigned)!(fabs(x) < 512.0);
    const double t = fma(x, HB_EXP_C0, 6755399441055744.0);   // 1.5 * 2^52: the low word of t is k = round(64 x / ln 2), two's complement
    const int k = __double2loint(t);
    double T;
#ifdef HB_HOST_EMU
    T = hb_kTab.x2[k & 63];
#else
    unsigned addr;
    asm("{\n\t.reg .u32 j;\n\tand.b32 j, %1, 63;\n\tmad.lo.u32 %0, j, 8, %2;\n\t}" : "=r"(addr) : "r"(k), "r"(cx.tab_s + (unsigned)(sizeof(double2) * HB_SC_N)));
    asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(addr));
#endif
    const double kf = t - 6755399441055744.0;
    double r = fma(-kf, HB_EXP_L1, x);
    r = fma(-kf, HB_EXP_L2, r);
    const double z = r * r;
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(r, q, 1.0 / 6.0);
    q = fma(r, q, 0.5);
    const double p = fma(z, q, r);
    const double v = fma(T, p, T);
    return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));   // * 2^m
  }
}

// hb_rcp: 1/d for the mass-matrix pivots.  MUFU.RCP64H seed (relative error <= 2^-20, measured over 6e9 operands:
// profiles/r1c/exp_rcp.txt) + ONE cubically convergent step x (1 + e + e^2), e = 1 - d x: 3 dependent DFMA, result
// within 1 ulp of IEEE 1/d (two Newton steps: 4 DFMA, correctly rounded — not worth a DFMA per RHS).  No
// denormal/overflow slow path: a pivot that needs one is flagged HB_FLAG_NOT_SPD / NONFINITE anyway.
HB_DEV double hb_rcp(double d) {
  double x;
#ifdef HB_HOST_EMU
  x = hb_emu_rcp64h(d);   // what MUFU.RCP64H delivers: 1/d to about 20 bits
#else
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
#endif
  const double e = fma(-d, x, 1.0);
  return fma(x, fma(e, e, e), x);
}

// hb_recip<FAST>: the `recip` / `/` of user tapes (e.g. the two-body potential -m1 m2 / r).  Fast path = hb_rcp for
// arguments whose exponent keeps the Newton iteration in range (2^-1000 < |x| < 2^1000); anything else (zero,
// denormal, huge, inf, nan) is redone on the IEEE-exact slow path.
template <bool FAST>
HB_DEV double hb_recip(HbCtx& cx, double x) {
  if constexpr (!FAST) {
    return 1.0 / x;
  } else {
    const unsigned e = ((unsigned)__double2hiint(x) >> 20) & 0x7ff;
    cx.oob |= (unsigned)(e - 23u > 2000u);   // biased exponent outside [23, 2023]
    return hb_rcp(x);
  }
}

// Compile-time loop: f(HbIdx<B>{}), ..., f(HbIdx<E-1>{}).  The Sys index tables are queried with
// true constant expressions (`if constexpr`), so structural zeros cost nothing — not even
// front-end time — and the emitted code contains only the surviving FMAs.
template <int I> struct HbIdx { static constexpr int value = I; };
template <int B, int E, class F>
HB_DEV void hb_static_for(F&& f) {
  if constexpr (E - B == 1) f(HbIdx<B>{});
  else if constexpr (E - B > 1) { hb_static_for<B, (B + E) / 2>(f); hb_static_for<(B + E) / 2, E>(f); }
}
#define HB_IDX(name, tag) constexpr int name = decltype(tag)::value

// ------------------------------------------------------------------ tiny dense algebra -----
HB_DEV constexpr int hb_tri(int j, int k) { return j * (j + 1) / 2 + k; }   // packed lower-triangular index

// A = J^T diag(w) J, lower triangle packed, skipping structurally-zero products.
// (reference: jmj = trj <> mm <> j, src/Numeric/Hamilton.hs:380)
template <class S>
HB_DEV void hb_mass(const double* wJ, const double* Jv, double* A) {
  hb_static_for<0, S::N>([&](auto jt) {
    HB_IDX(j, jt);
    hb_static_for<0, j + 1>([&](auto kt) {
      HB_IDX(k, kt);
      double acc = 0.0;
      hb_static_for<0, S::M>([&](auto it) {
        HB_IDX(i, it);
        if constexpr (S::jidx(i, j) >= 0 && S::jidx(i, k) >= 0) acc = fma(wJ[S::jidx(i, j)], Jv[S::jidx(i, k)], acc);
      });
      A[hb_tri(j, k)] = acc;
    });
  });
}

#ifndef HB_SOLVE_COLS
#define HB_SOLVE_COLS 1   // LDL^T and substitutions in column (axpy) order; 0: row (dot-product) order
#endif
// (A right-looking ordering of the LDL^T updates — runs of DFMAs sharing their first operand, for the operand-reuse cache —
// measured no different on the 12 x 12 chain: profiles/r2l/ab_chain12_ldlt.txt.)
// Solve A x = b for the packed SPD mass matrix; A is destroyed.  Replaces the reference's explicit
// `inv jmj` (src/Numeric/Hamilton.hs:381); a non-positive pivot is the analogue of hmatrix's
// singular-matrix exception.  N = 1, 2: closed form with ONE reciprocal (shortest dependency chain);
// N >= 3: LDL^T — no square roots, N reciprocals — then forward/diagonal/backward substitution.
// pivot test on the integer pipe: true unless d is a positive normal number (NaN passes here and is
// caught by the non-finite check on the result)
HB_DEV bool hb_bad_pivot(double d) { return __double2hiint(d) < 0x00100000; }
// The test is deferred: the solves only keep the smallest pivot high word seen (ONE integer min per pivot instead of a
// compare + select into the flag word); hb_ctx_flags applies hb_bad_pivot to it once per trajectory.
HB_DEV void hb_piv(int& minpiv, double d) { const int h = __double2hiint(d); minpiv = h < minpiv ? h : minpiv; }

struct HbRegVec {   // right-hand side held in registers
  const double* b;
  HB_DEV double operator()(int j) const { return b[j]; }
};
template <int STRIDE>
struct HbSmemVec {  // right-hand side parked in shared memory (element j at b[j * STRIDE]), read at the point of use
  const double* b;
  HB_DEV double operator()(int j) const { return hb_lds(b + j * STRIDE); }
};
// CLOSED: use the closed forms for N <= 3 (fast path).  Their pivot test looks at the determinant, which under- or overflows
// for inertias around 1e-160 / 1e+160 although the matrix is perfectly invertible (LAPACK's `inv` in the reference has no
// such problem): a failed test on the fast path therefore only sends the trajectory to the out-of-line slow path
// (hb_retry), which runs the scale-safe LDL^T for every N and raises HB_FLAG_NOT_SPD only if a pivot really is <= 0.
template <int N, bool CLOSED, class BV>
HB_DEV void hb_spd_solve_v(double* A, const BV b, double* x, int& minpiv);
template <int N, bool CLOSED>
HB_DEV void hb_spd_solve(double* A, const double* b, double* x, int& minpiv) { hb_spd_solve_v<N, CLOSED>(A, HbRegVec{b}, x, minpiv); }
template <int N, bool CLOSED, class BV>
HB_DEV void hb_spd_solve_v(double* A, const BV b, double* x, int& minpiv) {
  if constexpr (N == 1) {
    hb_piv(minpiv, A[0]);
    x[0] = b(0) * hb_rcp(A[0]);
  } else if constexpr (N == 2 && CLOSED) {
    const double det = fma(-A[1], A[1], A[0] * A[2]);   // diagonal product on its own: a literal when the system compiler emitted constant diagonals (pendulums)
    hb_piv(minpiv, A[0]);
    hb_piv(minpiv, det);
    const double id = hb_rcp(det);
    const double n0 = fma(A[2], b(0), -A[1] * b(1));
    const double n1 = fma(A[0], b(1), -A[1] * b(0));
    x[0] = n0 * id;
    x[1] = n1 * id;
#ifndef HB_SPD3_LDLT
  } else if constexpr (N == 3 && CLOSED) {
    // adjugate form: 6 cofactors, ONE reciprocal — a dependency chain of ~12 DFMA instead of LDL^T's three
    // back-to-back reciprocal chains (this engine runs at 4-6 warps per scheduler, so latency is throughput);
    // positive leading minors (a, ad - b^2, det) <=> SPD
    const double m00 = A[0], m10 = A[1], m11 = A[2], m20 = A[3], m21 = A[4], m22 = A[5];
    // products of two diagonal entries stand alone: literals when the system compiler emitted constant diagonals (pendulums)
    const double c00 = fma(-m21, m21, m11 * m22), c01 = fma(m20, m21, -m10 * m22), c02 = fma(m10, m21, -m20 * m11);
    const double c11 = fma(-m20, m20, m00 * m22), c12 = fma(m10, m20, -m00 * m21), c22 = fma(-m10, m10, m00 * m11);
    const double det = fma(m00, c00, fma(m10, c01, m20 * c02));
    hb_piv(minpiv, m00);
    hb_piv(minpiv, c22);
    hb_piv(minpiv, det);
    const double id = hb_rcp(det);
    const double b0 = b(0), b1 = b(1), b2 = b(2);
    x[0] = fma(c00, b0, fma(c01, b1, c02 * b2)) * id;
    x[1] = fma(c01, b0, fma(c11, b1, c12 * b2)) * id;
    x[2] = fma(c02, b0, fma(c12, b1, c22 * b2)) * id;
#endif
  } else {
    // LDL^T, fully unrolled by compile-time recursion (a rolled loop would index A dynamically and push it to local memory)
    double invd[N];
#if HB_SOLVE_COLS
    // Column-oriented ("right-looking", axpy) order: as soon as column k is final every element it touches is updated, so
    // consecutive DFMAs write DIFFERENT accumulators and share one operand (operand-reuse cache).  The row-oriented (dot
    // product) order below has the same dependency graph per element, but ptxas emits it as written: runs of DFMAs on ONE
    // accumulator, 8 clocks apart — with two warps per scheduler (large systems) a third of all warp samples of the chain's
    // step waited on those chains (profiles/r2u).  Per element the operations are those of the row-oriented form in the same
    // order, except that L_jk d_k is taken as the unscaled entry it came from and that the backward substitution accumulates in
    // descending instead of ascending column order (last-bit differences, on the accurate side).
    hb_static_for<0, N>([&](auto kt) {
      HB_IDX(k, kt);
      const double d = A[hb_tri(k, k)];
      hb_piv(minpiv, d);
      const double id = hb_rcp(d);
      invd[k] = id;
      double v[N - k > 1 ? N - k - 1 : 1];   // v_j = L_jk d_k = the unscaled column entry t_jk itself: no multiply (the row form
                                             // stores only L and recomputes it as (t_jk / d_k) d_k: 66 DMULs per 12 x 12 solve)
      hb_static_for<k + 1, N>([&](auto it) { HB_IDX(i, it); v[i - k - 1] = A[hb_tri(i, k)]; A[hb_tri(i, k)] = v[i - k - 1] * id; });   // t_ik -> L_ik
      hb_static_for<k + 1, N>([&](auto jt) {   // column j of the trailing matrix: A_ij -= L_ik t_jk, i >= j
        HB_IDX(j, jt);
        hb_static_for<j, N>([&](auto it) { HB_IDX(i, it); A[hb_tri(i, j)] = fma(-A[hb_tri(i, k)], v[j - k - 1], A[hb_tri(i, j)]); });
      });
    });
#else
    hb_static_for<0, N>([&](auto jt) {
      HB_IDX(j, jt);
      double v[j > 0 ? j : 1];
      double d = A[hb_tri(j, j)];
      hb_static_for<0, j>([&](auto kt) {
        HB_IDX(k, kt);
        v[k] = A[hb_tri(j, k)] * A[hb_tri(k, k)];   // L_jk d_k   (A_kk holds d_k)
        d = fma(-A[hb_tri(j, k)], v[k], d);
      });
      hb_piv(minpiv, d);
      A[hb_tri(j, j)] = d;
      const double id = hb_rcp(d);
      invd[j] = id;
      hb_static_for<j + 1, N>([&](auto it) {
        HB_IDX(i, it);
        double t = A[hb_tri(i, j)];
        hb_static_for<0, j>([&](auto kt) { HB_IDX(k, kt); t = fma(-A[hb_tri(i, k)], v[k], t); });
        A[hb_tri(i, j)] = t * id;
      });
    });
#endif
#if HB_SOLVE_COLS
#pragma unroll
    for (int j = 0; j < N; j++) x[j] = b(j);
    hb_static_for<0, N>([&](auto jt) {   // L y = b: y_j is final, every later row takes its share of it
      HB_IDX(j, jt);
      hb_static_for<j + 1, N>([&](auto it) { HB_IDX(i, it); x[i] = fma(-A[hb_tri(i, j)], x[j], x[i]); });
    });
#pragma unroll
    for (int j = 0; j < N; j++) x[j] *= invd[j];   // D z = y
    hb_static_for<0, N>([&](auto rt) {   // L^T x = z
      HB_IDX(r, rt);
      constexpr int j = N - 1 - r;
      hb_static_for<0, j>([&](auto it) { HB_IDX(i, it); x[i] = fma(-A[hb_tri(j, i)], x[j], x[i]); });
    });
#else
    hb_static_for<0, N>([&](auto jt) {   // L y = b
      HB_IDX(j, jt);
      double t = b(j);
      hb_static_for<0, j>([&](auto kt) { HB_IDX(k, kt); t = fma(-A[hb_tri(j, k)], x[k], t); });
      x[j] = t;
    });
#pragma unroll
    for (int j = 0; j < N; j++) x[j] *= invd[j];   // D z = y
    hb_static_for<0, N>([&](auto rt) {   // L^T x = z
      HB_IDX(r, rt);
      constexpr int j = N - 1 - r;
      double t = x[j];
      hb_static_for<j + 1, N>([&](auto kt) { HB_IDX(k, kt); t = fma(-A[hb_tri(k, j)], x[k], t); });
      x[j] = t;
    });
#endif
  }
}

// wJ[e] = w[row(e)] * J[e]
template <class S>
HB_DEV void hb_weigh(const double* w, const double* Jv, double* wJ) {
  hb_static_for<0, S::NJ>([&](auto et) { HB_IDX(e, et); wJ[e] = w[S::jrow(e)] * Jv[e]; });
}
// out[j] = sum_i A_e[i][j] * x[i]   (J^T-type product over the packed non-zeros)
template <class S>
HB_DEV void hb_jt_mul(const double* Je, const double* x, double* out) {
#pragma unroll
  for (int j = 0; j < S::N; j++) out[j] = 0.0;
  hb_static_for<0, S::NJ>([&](auto et) { HB_IDX(e, et); out[S::jcol(e)] = fma(Je[e], x[S::jrow(e)], out[S::jcol(e)]); });
}
// out[i] = sum_j A_e[i][j] * v[j]
template <class S>
HB_DEV void hb_j_mul(const double* Je, const double* v, double* out) {
#pragma unroll
  for (int i = 0; i < S::M; i++) out[i] = 0.0;
  hb_static_for<0, S::NJ>([&](auto et) { HB_IDX(e, et); out[S::jrow(e)] = fma(Je[e], v[S::jcol(e)], out[S::jrow(e)]); });
}

// Symbolic form of hamEqs: hpre -> SPD solve -> hpost.  For large systems the values carried from hpre to hpost are
// parked in dynamic shared memory across the solve (hb_dsm + 3 * 2n * B doubles, column per thread).
template <class S, bool FAST, class PV>
HB_DEV void hb_ham_eqs_sym(HbCtx& cx, const double* prm, const double* qq, const PV p, double* dq, double* dp, int& flag) {
  constexpr int N = S::N, NE = S::NE;
  double A[N * (N + 1) / 2];
  if constexpr (N >= HB_BIG_N && NE > 0) {
    constexpr int B = HB_BLOCK_OF(N);
    double* es = hb_dsm + 3 * 2 * N * B + threadIdx.x;
    {
      double E[NE];
      S::template hpre<FAST>(cx, prm, qq, A, E);
#pragma unroll
      for (int e = 0; e < NE; e++) hb_sts(es + e * B, E[e]);
    }
    hb_spd_solve_v<N, FAST>(A, p, dq, cx.minpiv);
    double E[NE];
#pragma unroll
    for (int e = 0; e < NE; e++) E[e] = hb_lds(es + e * B);
    S::template hpost<FAST>(cx, prm, qq, E, dq, dp);
  } else {
    double E[NE > 0 ? NE : 1];
    S::template hpre<FAST>(cx, prm, qq, A, E);
    hb_spd_solve_v<N, FAST>(A, p, dq, cx.minpiv);
    S::template hpost<FAST>(cx, prm, qq, E, dq, dp);
  }
}

// --------------------------------------------------------------------------- hamEqs --------
// (dq, dp) = hamEqs(q, p)  (src/Numeric/Hamilton.hs:370-387):
//   dq   = M^-1 p                                                   (:386)
//   dp_j = +(p . M^-1 J^T W H_j M^-1 p) - dU/dq_j                   (:382-387, sign from -dHdq :375)
// evaluated as  a = W J dq,  dp_j = sum_i a_i (H_j dq)_i - gU_j  — one SPD solve instead of the
// reference's explicit inverse and 5n mat-vecs, and H_j never materialised beyond its non-zeros.
template <class S, bool FAST>
HB_DEV void hb_ham_eqs(HbCtx& cx, const double* prm, const double* w, const double* q, const double* p,
                       double* dq, double* dp, int& flag) {
  constexpr int N = S::N, M = S::M, NJ = S::NJ, NH = S::NH;
  double qq[N];
#pragma unroll
  for (int j = 0; j < N; j++) qq[j] = q[j];
  if constexpr (S::SYMH) {
    // the system compiler resolved M(q) and 1/2 v^T dM/dq_j v - dU/dq_j symbolically (csrc/polyform.cpp): only the
    // entries that survive the Pythagorean cancellations are evaluated; the SPD solve in between is all that is left here
    (void)w;
    hb_ham_eqs_sym<S, FAST>(cx, prm, qq, HbRegVec{p}, dq, dp, flag);
    return;
  }
  double Jv[NJ > 0 ? NJ : 1], Hv[NH > 0 ? NH : 1], gU[N];
  S::template derivs<FAST>(cx, prm, qq, Jv, Hv, gU);
  double wJ[NJ > 0 ? NJ : 1];
  hb_weigh<S>(w, Jv, wJ);
  double A[N * (N + 1) / 2];
  hb_mass<S>(wJ, Jv, A);
  hb_spd_solve<N, FAST>(A, p, dq, cx.minpiv);
  double a[M];
  hb_j_mul<S>(wJ, dq, a);
#pragma unroll
  for (int j = 0; j < N; j++) dp[j] = -gU[j];
  // s_g = sum_i a_i H_i,(j,k) over the entries of group g = (j, k); then dp_j += s_g v_k (and dp_k += s_g v_j)
  constexpr int NG = S::NG;
  double sg[NG > 0 ? NG : 1];
#pragma unroll
  for (int g = 0; g < NG; g++) sg[g] = 0.0;
  hb_static_for<0, NH>([&](auto et) { HB_IDX(e, et); sg[S::hgrp(e)] = fma(a[S::hrow(e)], Hv[e], sg[S::hgrp(e)]); });
  hb_static_for<0, NG>([&](auto gt) {
    HB_IDX(g, gt);
    dp[S::gj(g)] = fma(sg[g], dq[S::gk(g)], dp[S::gj(g)]);
    if constexpr (S::gj(g) != S::gk(g)) dp[S::gk(g)] = fma(sg[g], dq[S::gj(g)], dp[S::gk(g)]);
  });
}
// F(y) on the packed Phase vector y = [q, p]  (fromPs/toPs, src/Numeric/Hamilton.hs:457-462)
template <class S, bool FAST>
HB_DEV void hb_rhs(HbCtx& cx, const double* prm, const double* w, const double* y, double* dy, int& flag) {
  hb_ham_eqs<S, FAST>(cx, prm, w, y, y + S::N, dy, dy + S::N, flag);
}

// momenta (src/Numeric/Hamilton.hs:262-269): p = J^T (W (J v))
template <class S, bool FAST>
HB_DEV void hb_momenta(HbCtx& cx, const double* prm, const double* w, const double* q, const double* v, double* p) {
  constexpr int N = S::N, M = S::M, NJ = S::NJ;
  double Jv[NJ > 0 ? NJ : 1], qq[N], t[M];
#pragma unroll
  for (int j = 0; j < N; j++) qq[j] = q[j];
  if constexpr (S::SYMH) {   // p = A(q) v with the symbolic packed mass matrix
    (void)w;
    double A[N * (N + 1) / 2];
    S::template smass<FAST>(cx, prm, qq, A);
#pragma unroll
    for (int j = 0; j < N; j++) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; k++) acc = fma(A[j >= k ? hb_tri(j, k) : hb_tri(k, j)], v[k], acc);
      p[j] = acc;
    }
    return;
  }
  S::template jac<FAST>(cx, prm, qq, Jv);
  hb_j_mul<S>(Jv, v, t);
#pragma unroll
  for (int i = 0; i < M; i++) t[i] *= w[i];
  hb_jt_mul<S>(Jv, t, p);
}
// velocities (src/Numeric/Hamilton.hs:316-324): v = (J^T W J)^-1 p ; optionally also U(q)
template <class S, bool FAST, bool WITH_U>
HB_DEV void hb_velocities(HbCtx& cx, const double* prm, const double* w, const double* q, const double* p, double* v,
                          double& U, int& flag) {
  constexpr int N = S::N, NJ = S::NJ;
  double Jv[NJ > 0 ? NJ : 1], wJ[NJ > 0 ? NJ : 1], qq[N];
#pragma unroll
  for (int j = 0; j < N; j++) qq[j] = q[j];
  if constexpr (S::SYMH) {
    (void)w;
    double A[N * (N + 1) / 2];
    if constexpr (WITH_U) S::template smass_pot<FAST>(cx, prm, qq, A, U); else S::template smass<FAST>(cx, prm, qq, A);
    hb_spd_solve<N, FAST>(A, p, v, cx.minpiv);
    return;
  }
  if constexpr (WITH_U) S::template jac_pot<FAST>(cx, prm, qq, Jv, U); else S::template jac<FAST>(cx, prm, qq, Jv);
  hb_weigh<S>(w, Jv, wJ);
  double A[N * (N + 1) / 2];
  hb_mass<S>(wJ, Jv, A);
  hb_spd_solve<N, FAST>(A, p, v, cx.minpiv);
}

// ------------------------------------------------------------------------ integrators ------
// Classical RK4 — the unit of BASELINE.json's metric (4 hamEqs evaluations per step).
template <class S, bool FAST>
HB_DEV void hb_rk4_step(HbCtx& cx, const double* prm, const double* w, double (&y)[2 * S::N], double dt, double h6, double hh, int& flag) {
  constexpr int D = 2 * S::N;
  double k[D], yt[D], acc[D];
  hb_rhs<S, FAST>(cx, prm, w, y, k, flag);
#pragma unroll
  for (int c = 0; c < D; c++) { acc[c] = k[c]; yt[c] = fma(hh, k[c], y[c]); }
  hb_rhs<S, FAST>(cx, prm, w, yt, k, flag);
#pragma unroll
  for (int c = 0; c < D; c++) { acc[c] = fma(2.0, k[c], acc[c]); yt[c] = fma(hh, k[c], y[c]); }
  hb_rhs<S, FAST>(cx, prm, w, yt, k, flag);
#pragma unroll
  for (int c = 0; c < D; c++) { acc[c] = fma(2.0, k[c], acc[c]); yt[c] = fma(dt, k[c], y[c]); }
  hb_rhs<S, FAST>(cx, prm, w, yt, k, flag);
#pragma unroll
  for (int c = 0; c < D; c++) y[c] = fma(h6, acc[c] + k[c], y[c]);
}

// Large systems (2n >= 16): y, acc and yt live in shared memory (column per thread: conflict-free), so the
// registers are left to the RHS — a 12x12 LDL^T alone keeps ~100 doubles live.
template <int D>
HB_DEV double* hb_rk_smem() { return hb_dsm; }   // 3 * D * HB_BLOCK_OF(D / 2) doubles at the start of the dynamic window
// F(y) with y read from a shared-memory column (element c at src[c * B]): q goes to registers, p is read where the
// forward substitution needs it, so nothing but the mass matrix is live across the factorisation.
template <class S, bool FAST>
HB_DEV void hb_rhs_sm(HbCtx& cx, const double* prm, const double* w, const double* src, double* k, int& flag) {
  constexpr int N = S::N, B = HB_BLOCK_OF(S::N);
  if constexpr (S::SYMH) {
    (void)w;
    double qq[N];
#pragma unroll
    for (int j = 0; j < N; j++) qq[j] = src[j * B];
    hb_ham_eqs_sym<S, FAST>(cx, prm, qq, HbSmemVec<B>{src + N * B}, k, k + N, flag);
  } else {
    double in[2 * N];
#pragma unroll
    for (int c = 0; c < 2 * N; c++) in[c] = src[c * B];
    hb_rhs<S, FAST>(cx, prm, w, in, k, flag);
  }
}
#ifndef HB_RK4_STAGE_LOOP
#define HB_RK4_STAGE_LOOP 1   // large systems: the four stages are ONE loop around ONE copy of the RHS (instruction cache, see below)
#endif
// The RHS of a large system is thousands of straight-line instructions (chain (24,12): 1,600 = 26 KB).  Inlined four times the
// step body is 104 KB — more than the 32 KB instruction cache level next to the SM can hold — and with two warps per
// scheduler at different places of it every warp waits for instruction fetches (ncu: stall_no_instruction 1.09 cycles per
// issue, profiles/r2l).  The stages therefore run as a loop (never unrolled) whose epilogue is selected by the
// warp-uniform stage index; every sum is the same fma as in the unrolled form (fma(1, k, 0) == k, fma(1, k, acc) == acc + k),
// so the results are bit-identical to it.
template <class S, bool FAST>
HB_DEV void hb_rk4_step_sm(HbCtx& cx, const double* prm, const double* w, double* sm, double dt, double h6, int& flag) {
  constexpr int D = 2 * S::N, B = HB_BLOCK_OF(S::N);
  double* ys = sm + threadIdx.x;            // y[c]   = ys[c * B]
  double* as = ys + D * B;                  // acc[c]
  double* ts = as + D * B;                  // yt[c]
  const double hh = 0.5 * dt;
#if HB_RK4_STAGE_LOOP
#pragma unroll 1
  for (int st = 0; st < 4; st++) {
    double k[D];
    hb_rhs_sm<S, FAST>(cx, prm, w, st == 0 ? ys : ts, k, flag);
    if (st == 0) {
#pragma unroll
      for (int c = 0; c < D; c++) { as[c * B] = k[c]; ts[c * B] = fma(hh, k[c], ys[c * B]); }
    } else if (st < 3) {
      const double cs = st == 1 ? hh : dt;
#pragma unroll
      for (int c = 0; c < D; c++) { as[c * B] = fma(2.0, k[c], as[c * B]); ts[c * B] = fma(cs, k[c], ys[c * B]); }
    } else {
#pragma unroll
      for (int c = 0; c < D; c++) ys[c * B] = fma(h6, as[c * B] + k[c], ys[c * B]);
    }
  }
#else
  double k[D];
  hb_rhs_sm<S, FAST>(cx, prm, w, ys, k, flag);
#pragma unroll
  for (int c = 0; c < D; c++) { as[c * B] = k[c]; ts[c * B] = fma(hh, k[c], ys[c * B]); }
  hb_rhs_sm<S, FAST>(cx, prm, w, ts, k, flag);
#pragma unroll
  for (int c = 0; c < D; c++) { as[c * B] = fma(2.0, k[c], as[c * B]); ts[c * B] = fma(hh, k[c], ys[c * B]); }
  hb_rhs_sm<S, FAST>(cx, prm, w, ts, k, flag);
#pragma unroll
  for (int c = 0; c < D; c++) { as[c * B] = fma(2.0, k[c], as[c * B]); ts[c * B] = fma(dt, k[c], ys[c * B]); }
  hb_rhs_sm<S, FAST>(cx, prm, w, ts, k, flag);
#pragma unroll
  for (int c = 0; c < D; c++) ys[c * B] = fma(h6, as[c * B] + k[c], ys[c * B]);
#endif
}

// GSL-semantics adaptive RKF45 — what the reference's stepHam/evolveHam actually run
// (`odeSolveV RKf45 hi eps eps`, src/Numeric/Hamilton.hs:445-448): GSL 2.x rkf45 stepper (5th-order
// solution advanced, 4th/5th difference as error), standard controller with a_y = a_dydt = 1 and
// gsl_odeiv2_evolve_apply's accept/reject/FSAL logic.  One thread runs its own controller, so step
// sequences may differ between neighbouring trajectories (warp divergence only in trip counts).
struct HbRkf45 {
  // Fehlberg tableau as in GSL rkf45.c
  static constexpr double ah0 = 1.0 / 4.0;
  static constexpr double b30 = 3.0 / 32.0, b31 = 9.0 / 32.0;
  static constexpr double b40 = 1932.0 / 2197.0, b41 = -7200.0 / 2197.0, b42 = 7296.0 / 2197.0;
  static constexpr double b50 = 8341.0 / 4104.0, b51 = -32832.0 / 4104.0, b52 = 29440.0 / 4104.0, b53 = -845.0 / 4104.0;
  static constexpr double b60 = -6080.0 / 20520.0, b61 = 41040.0 / 20520.0, b62 = -28352.0 / 20520.0,
                          b63 = 9295.0 / 20520.0, b64 = -5643.0 / 20520.0;
  static constexpr double c1 = 902880.0 / 7618050.0, c3 = 3953664.0 / 7618050.0, c4 = 3855735.0 / 7618050.0,
                          c5 = -1371249.0 / 7618050.0, c6 = 277020.0 / 7618050.0;
  static constexpr double e1 = 1.0 / 360.0, e3 = -128.0 / 4275.0, e4 = -2197.0 / 75240.0, e5 = 1.0 / 50.0, e6 = 2.0 / 55.0;
  static constexpr double eps = 1.49012e-08;   // src/Numeric/Hamilton.hs:448
};

// One rkf45_apply: y <- y + h * (5th-order increment); yerr; dydt_out = F(y_new).  k1 = dydt_in.
// Every product and sum is written as an explicit fma / multiply: no expression is left for the compiler to contract one
// way or another, so all instances of this code (layout-specialised, ahead-of-time, NVRTC) give bit-identical results.
template <class S, bool FAST>
HB_DEV void hb_rkf45_apply(HbCtx& cx, const double* prm, const double* w, double h, double (&y)[2 * S::N],
                           const double (&k1)[2 * S::N], double (&yerr)[2 * S::N],
                           double (&dydt_out)[2 * S::N], int& flag) {
  constexpr int D = 2 * S::N;
  typedef HbRkf45 T;
  const double h4 = T::ah0 * h;
  if constexpr (D >= 16 && HB_RK4_STAGE_LOOP) {
    // Large systems: ONE copy of the RHS inside a loop over the six evaluations (instruction cache, see hb_rk4_step_sm); the
    // stage derivatives are addressed by the loop index, i.e. they live in local memory — where ptxas spills them anyway
    // when the RHS needs every register.  Each case is the same fma chain as the straight-line form below: bit-identical.
    double K[6][D], yt[D];   // K[0..4] = k2..k6, K[5] = dydt_out
#pragma unroll 1
    for (int st = 0; st < 6; st++) {
      if (st == 0) {
#pragma unroll
        for (int c = 0; c < D; c++) yt[c] = fma(h4, k1[c], y[c]);
      } else if (st == 1) {
#pragma unroll
        for (int c = 0; c < D; c++) yt[c] = fma(h, fma(T::b30, k1[c], T::b31 * K[0][c]), y[c]);
      } else if (st == 2) {
#pragma unroll
        for (int c = 0; c < D; c++) yt[c] = fma(h, fma(T::b40, k1[c], fma(T::b41, K[0][c], T::b42 * K[1][c])), y[c]);
      } else if (st == 3) {
#pragma unroll
        for (int c = 0; c < D; c++) yt[c] = fma(h, fma(T::b50, k1[c], fma(T::b51, K[0][c], fma(T::b52, K[1][c], T::b53 * K[2][c]))), y[c]);
      } else if (st == 4) {
#pragma unroll
        for (int c = 0; c < D; c++)
          yt[c] = fma(h, fma(T::b60, k1[c], fma(T::b61, K[0][c], fma(T::b62, K[1][c], fma(T::b63, K[2][c], T::b64 * K[3][c])))), y[c]);
      } else {
#pragma unroll
        for (int c = 0; c < D; c++) {
          const double di = fma(T::c1, k1[c], fma(T::c3, K[1][c], fma(T::c4, K[2][c], fma(T::c5, K[3][c], T::c6 * K[4][c]))));
          y[c] = fma(h, di, y[c]);
          yerr[c] = h * fma(T::e1, k1[c], fma(T::e3, K[1][c], fma(T::e4, K[2][c], fma(T::e5, K[3][c], T::e6 * K[4][c]))));
          yt[c] = y[c];
        }
      }
      hb_rhs<S, FAST>(cx, prm, w, yt, K[st], flag);
    }
#pragma unroll
    for (int c = 0; c < D; c++) dydt_out[c] = K[5][c];
    return;
  }
  double k2[D], k3[D], k4[D], k5[D], k6[D], yt[D];
#pragma unroll
  for (int c = 0; c < D; c++) yt[c] = fma(h4, k1[c], y[c]);
  hb_rhs<S, FAST>(cx, prm, w, yt, k2, flag);
#pragma unroll
  for (int c = 0; c < D; c++) yt[c] = fma(h, fma(T::b30, k1[c], T::b31 * k2[c]), y[c]);
  hb_rhs<S, FAST>(cx, prm, w, yt, k3, flag);
#pragma unroll
  for (int c = 0; c < D; c++) yt[c] = fma(h, fma(T::b40, k1[c], fma(T::b41, k2[c], T::b42 * k3[c])), y[c]);
  hb_rhs<S, FAST>(cx, prm, w, yt, k4, flag);
#pragma unroll
  for (int c = 0; c < D; c++) yt[c] = fma(h, fma(T::b50, k1[c], fma(T::b51, k2[c], fma(T::b52, k3[c], T::b53 * k4[c]))), y[c]);
  hb_rhs<S, FAST>(cx, prm, w, yt, k5, flag);
#pragma unroll
  for (int c = 0; c < D; c++)
    yt[c] = fma(h, fma(T::b60, k1[c], fma(T::b61, k2[c], fma(T::b62, k3[c], fma(T::b63, k4[c], T::b64 * k5[c])))), y[c]);
  hb_rhs<S, FAST>(cx, prm, w, yt, k6, flag);
#pragma unroll
  for (int c = 0; c < D; c++) {
    const double di = fma(T::c1, k1[c], fma(T::c3, k3[c], fma(T::c4, k4[c], fma(T::c5, k5[c], T::c6 * k6[c]))));
    y[c] = fma(h, di, y[c]);
    yerr[c] = h * fma(T::e1, k1[c], fma(T::e3, k3[c], fma(T::e4, k4[c], fma(T::e5, k5[c], T::e6 * k6[c]))));
  }
  hb_rhs<S, FAST>(cx, prm, w, y, dydt_out, flag);
}

// Per-trajectory evolve state carried across output times like hmatrix-gsl's loop carries
// `e` and `h` (primed == false  <=>  e->count == 0: dydt_out not yet valid).
template <int D>
struct HbEvolve {
  double h;
  double dydt[D];   // dydt_out of the last accepted step (FSAL)
  bool primed;
};

// Integrate y from t to t1 with gsl_odeiv2_evolve_apply semantics: `while (t < t1) evolve_apply`.
template <class S, bool FAST>
HB_DEV void hb_rkf45_to(HbCtx& cx, const double* prm, const double* w, double (&y)[2 * S::N], double& t, double t1,
                        HbEvolve<2 * S::N>& e, int& flag) {
  constexpr int D = 2 * S::N;
  typedef HbRkf45 T;
  int guard = 0, attempts = 0;
  while (t < t1) {
    const double t0 = t;
    double h0 = e.h;
    const double dt = t1 - t0;
    double y0[D], k1[D], yerr[D], dout[D];
#pragma unroll
    for (int c = 0; c < D; c++) y0[c] = y[c];
    if (!e.primed) { hb_rhs<S, FAST>(cx, prm, w, y, k1, flag); e.primed = true; }
    else {
#pragma unroll
      for (int c = 0; c < D; c++) k1[c] = e.dydt[c];
    }
    bool final_step;
    for (;;) {   // try_step
      if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) { h0 = dt; final_step = true; } else final_step = false;
      hb_rkf45_apply<S, FAST>(cx, prm, w, h0, y, k1, yerr, dout, flag);
      t = final_step ? t1 : t0 + h0;
      // GSL (and with it the reference) would spin forever on a step that does not advance t (h = 0 after a zero-length first
      // grid interval, or h below the spacing of t); on a GPU that hangs the stream: give the trajectory up instead
      if ((!final_step && t == t0) || ++attempts > 20000000) { flag |= HB_FLAG_STEP_FAILED; e.h = h0; t = t1; return; }
      // std_control_hadjust, ord = 5
      const double h_old = h0;
      // Shortcut for the common case: every error ratio <= 3.3e-5 implies rmax < 0.5 and 0.9 rmax^(-1/6) >= 5.02, which
      // GSL clamps to exactly 5 — so the four divisions and the pow() of the general path (about a third of the FP64 work
      // of an attempt) are not needed to reproduce its decision bit for bit.  (0.9/5)^6 = 3.40e-5; the 3 % margin covers
      // the roundings of the test; NaNs fail the comparison and take the general path.)
      bool tiny = true;
#pragma unroll
      for (int c = 0; c < D; c++) {
        const double D0 = fma(T::eps, fabs(y[c]) + fabs(h_old * dout[c]), T::eps);
        tiny = tiny && (fabs(yerr[c]) <= 3.3e-5 * D0);
      }
#ifndef HB_NO_RKF45_SHORTCUT
      if (tiny) { h0 = 5.0 * h_old; break; }
#endif
      double rmax = 2.2250738585072014e-308;   // DBL_MIN
#pragma unroll
      for (int c = 0; c < D; c++) {
        const double D0 = fma(T::eps, fabs(y[c]) + fabs(h_old * dout[c]), T::eps);
        const double r = fabs(yerr[c]) / fabs(D0);
        rmax = r > rmax ? r : rmax;
      }
      const bool nonfinite = !hb_finite(rmax);
      if (rmax > 1.1) {
        double r = 0.9 / pow(rmax, 1.0 / 5.0);
        if (r < 0.2) r = 0.2;
        h0 = r * h_old;
        const double t_next = t + h0;
        if (fabs(h0) < fabs(h_old) && t_next != t && ++guard < 100000) {
#pragma unroll
          for (int c = 0; c < D; c++) y[c] = y0[c];   // undo, retry with the smaller h
          continue;
        }
        flag |= HB_FLAG_STEP_FAILED;   // GSL_FAILURE: cannot reach tolerance, cannot shrink h
        e.h = h0;
        t = t1;                        // give up on this trajectory (state is what the failed step produced)
        return;
      } else if (rmax < 0.5) {
        double r = 0.9 / pow(rmax, 1.0 / 6.0);
        if (r > 5.0) r = 5.0;
        if (r < 1.0) r = 1.0;
        h0 = r * h_old;
      }
      if (nonfinite) { flag |= HB_FLAG_NONFINITE; t = t1; return; }
      break;
    }
#pragma unroll
    for (int c = 0; c < D; c++) e.dydt[c] = dout[c];
    if (!final_step) e.h = h0;
  }
}

// ------------------------------------------------------------------ per-trajectory bodies ----
// LAY: layout known at compile time inside a kernel instance (-1 = read a.layout; 0 = array of Phases in device memory).
// The hot step kernels dispatch once per launch into a LAY = 0 instance, so the per-trajectory load/store carries no SOA
// and no warp-transposed-store code at all — ptxas otherwise if-converts the layout branches and the AOS path issues ~26
// predicated-off instructions per trajectory.  NST: 1 = exactly one step per launch, compiled as straight-line code (the
// loop over a.nsteps makes the state a loop-carried value and ptxas copies the prefetched Phase into it: 8-16 moves per
// trajectory); 0 = a.nsteps steps.
#ifndef HB_LAYSPEC
#define HB_LAYSPEC 1
#endif
#ifndef HB_REG_PREFETCH
#define HB_REG_PREFETCH 1  // plain-load path, records of <= 8 doubles: register prefetch of the next Phase
#endif
#ifndef HB_ASYNC_STAGE
#define HB_ASYNC_STAGE 1   // small systems, array of records: the next Phase is staged by cp.async into shared memory (0: plain loads)
#endif
template <int LAY> HB_DEV int hb_lay_of(const HbKArgs& a) { if constexpr (LAY < 0) return a.layout; else return LAY; }
// HB_FLAG_* bits of one trajectory: what the integrators raised + the deferred pivot test
HB_DEV int hb_ctx_flags(const HbCtx& cx, int flag) { return cx.minpiv < 0x00100000 ? (flag | HB_FLAG_NOT_SPD) : flag; }
// fast path only: redo this trajectory out of line (argument outside a fast primitive's domain, or a failed pivot test)
HB_DEV bool hb_retry(const HbCtx& cx) { return cx.oob != 0 || cx.minpiv < 0x00100000; }
// Each hb_traj_* processes trajectory i completely (compute -> store; the kernel body loads).  FAST=true is inlined
// into the kernel; if any fast primitive left its domain (cx.oob) nothing is stored and the kernel
// re-runs that one trajectory through the out-of-line FAST=false instance (HB_KERNEL_BODY below).
template <int D, class I>
HB_DEV void hb_finish(const HbKArgs& a, I i, const double (&y)[D], const HbCtx& cx, int flag) {
  if (a.flags == nullptr) return;   // (uniform) nobody asked: no finite checks, no flag word
  flag = hb_ctx_flags(cx, flag);
  bool ok = true;
#pragma unroll
  for (int c = 0; c < D; c++) ok = ok && hb_finite(y[c]);
  if (!ok) flag |= HB_FLAG_NONFINITE;
  if (flag) a.flags[i] |= flag;
}

// stepHam iterated with the fixed RK4 stepper
template <class S, bool FAST, int LAY, class I, int NST = 0>
HB_DEV void hb_traj_step_rk4(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N;
  double y[D];
  int flag = 0;
  if constexpr (D >= 16) {
    constexpr int B = HB_BLOCK_OF(S::N);
    double* sm = hb_rk_smem<D>();
#pragma unroll
    for (int c = 0; c < D; c++) sm[c * B + threadIdx.x] = yin[c];
    for (int s = 0; s < a.nsteps; s++) hb_rk4_step_sm<S, FAST>(cx, a.prm, w, sm, a.dt, a.dt6, flag);
#pragma unroll
    for (int c = 0; c < D; c++) y[c] = sm[c * B + threadIdx.x];
  } else {
    hb_copy<D>(yin, y);
    if constexpr (NST == 1) hb_rk4_step<S, FAST>(cx, a.prm, w, y, a.dt, a.dt6, a.dth, flag);
    else for (int s = 0; s < a.nsteps; s++) hb_rk4_step<S, FAST>(cx, a.prm, w, y, a.dt, a.dt6, a.dth, flag);
  }
  if (FAST && hb_retry(cx)) return;
  hb_store<D, I>(a.out, i, (I)a.N, hb_lay_of<LAY>(a), y, cx.xp);
  hb_finish<D, I>(a, i, y, cx, flag);
}
// stepHam iterated with reference semantics: each step is a fresh adaptive solve over (0, dt)
template <class S, bool FAST, int LAY, class I>
HB_DEV void hb_traj_step_rkf45(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N;
  double y[D];
  hb_copy<D>(yin, y);
  int flag = 0;
  if (a.dt > 0.0) {   // stepHam r with r <= 0 returns the Phase unchanged: hmatrix-gsl's `while (t < t1)` never runs
    for (int s = 0; s < a.nsteps; s++) {
      HbEvolve<D> e;
      e.h = a.dt / 100;   // hi = (t1 - t0)/100, src/Numeric/Hamilton.hs:447
      e.primed = false;
      double t = 0.0;
      hb_rkf45_to<S, FAST>(cx, a.prm, w, y, t, a.dt, e, flag);
    }
  }
  if (FAST && hb_retry(cx)) return;
  hb_store<D, I>(a.out, i, (I)a.N, hb_lay_of<LAY>(a), y, cx.xp);
  hb_finish<D, I>(a, i, y, cx, flag);
}
// evolveHam over a shared time grid; out[k] = batch at ts[k]
template <class S, bool FAST, int LAY, bool ADAPTIVE, class I>
HB_DEV void hb_traj_evolve(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N;
  double y[D];
  hb_copy<D>(yin, y);
  if (!FAST || !hb_retry(cx)) hb_store<D, I>(a.out, i, (I)a.N, hb_lay_of<LAY>(a), y, cx.xp);   // row 0 is the initial state
  int flag = 0;
  HbEvolve<D> e;
  e.h = (a.ts[1] - a.ts[0]) / 100;
  e.primed = false;
  double t = a.ts[0];
  for (int k = 1; k < a.s; k++) {
    const double tk = a.ts[k];
    if constexpr (ADAPTIVE) {
      hb_rkf45_to<S, FAST>(cx, a.prm, w, y, t, tk, e, flag);
    } else {
      const double h = (tk - t) / a.substeps;
      const double h6 = h / 6.0;
      if constexpr (D >= 16) {   // large systems: RK vectors in shared memory, stages as a loop (see hb_rk4_step_sm)
        constexpr int B = HB_BLOCK_OF(S::N);
        double* sm = hb_rk_smem<D>();
#pragma unroll
        for (int c = 0; c < D; c++) sm[c * B + threadIdx.x] = y[c];
        for (int s = 0; s < a.substeps; s++) hb_rk4_step_sm<S, FAST>(cx, a.prm, w, sm, h, h6, flag);
#pragma unroll
        for (int c = 0; c < D; c++) y[c] = sm[c * B + threadIdx.x];
      } else {
        for (int s = 0; s < a.substeps; s++) hb_rk4_step<S, FAST>(cx, a.prm, w, y, h, h6, 0.5 * h, flag);
      }
      t = tk;
    }
    if (FAST && hb_retry(cx)) return;   // rows written so far are rewritten by the slow retry (out never aliases in)
    hb_store<D, I>(a.out + (size_t)k * (size_t)a.N * D, i, (I)a.N, hb_lay_of<LAY>(a), y, cx.xp);
  }
  hb_finish<D, I>(a, i, y, cx, flag);
}
template <class S, bool FAST, int LAY, class I>
HB_DEV void hb_traj_evolve_rk4(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) { hb_traj_evolve<S, FAST, LAY, false, I>(a, i, yin, w, cx); }
template <class S, bool FAST, int LAY, class I>
HB_DEV void hb_traj_evolve_rkf45(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) { hb_traj_evolve<S, FAST, LAY, true, I>(a, i, yin, w, cx); }

template <class S, bool FAST, int LAY, class I>
HB_DEV void hb_traj_ham_eqs(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N;
  double y[D], dy[D];
  hb_copy<D>(yin, y);
  int flag = 0;
  hb_rhs<S, FAST>(cx, a.prm, w, y, dy, flag);
  if (FAST && hb_retry(cx)) return;
  hb_store<D, I>(a.out, i, (I)a.N, hb_lay_of<LAY>(a), dy, cx.xp);
  hb_finish<D, I>(a, i, dy, cx, flag);
}
template <class S, bool FAST, int LAY, class I>
HB_DEV void hb_traj_to_phase(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) {   // Config [q, v] -> Phase [q, p]
  constexpr int D = 2 * S::N, N = S::N;
  double c[D], y[D];
  hb_copy<D>(yin, c);
#pragma unroll
  for (int j = 0; j < N; j++) y[j] = c[j];
  hb_momenta<S, FAST>(cx, a.prm, w, c, c + N, y + N);
  if (FAST && hb_retry(cx)) return;
  hb_store<D, I>(a.out, i, (I)a.N, hb_lay_of<LAY>(a), y, cx.xp);
}
template <class S, bool FAST, int LAY, class I>
HB_DEV void hb_traj_from_phase(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) {   // Phase [q, p] -> Config [q, v]
  constexpr int D = 2 * S::N, N = S::N;
  double y[D], c[D];
  hb_copy<D>(yin, y);
  int flag = 0;
  double U;
#pragma unroll
  for (int j = 0; j < N; j++) c[j] = y[j];
  hb_velocities<S, FAST, false>(cx, a.prm, w, y, y + N, c + N, U, flag);
  if (FAST && hb_retry(cx)) return;
  hb_store<D, I>(a.out, i, (I)a.N, hb_lay_of<LAY>(a), c, cx.xp);
  hb_finish<D, I>(a, i, c, cx, flag);
}
// out4[i] = (keP, pe, hamiltonian, lagrangian)
template <class S, bool FAST, int LAY, class I>
HB_DEV void hb_traj_energies(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N, N = S::N;
  double y[D], v[N];
  hb_copy<D>(yin, y);
  int flag = 0;
  double U;
  hb_velocities<S, FAST, true>(cx, a.prm, w, y, y + N, v, U, flag);
  if (FAST && hb_retry(cx)) return;
  double T = 0.0;
#pragma unroll
  for (int j = 0; j < N; j++) T = fma(v[j], y[N + j], T);
  T *= 0.5;   // (vs <.> ps) / 2, src/Numeric/Hamilton.hs:349
  double o[4] = {T, U, T + U, T - U};
  hb_store<4, I>(a.out, i, (I)a.N, a.layout == 2 ? 2 : 0, o, cx.xp);
  hb_finish<4, I>(a, i, o, cx, flag);
}
template <class S, bool FAST, int LAY, class I>
HB_DEV void hb_traj_upos(const HbKArgs& a, I i, const double* yin, const double* w, HbCtx& cx) {   // underlyingPos
  constexpr int N = S::N, M = S::M;
  double q[N], x[M];
  (void)w;
  hb_copy<N>(yin, q);
  S::template pos<FAST>(cx, a.prm, q, x);
  if (FAST && hb_retry(cx)) return;
  hb_store<M, I>(a.out, i, (I)a.N, hb_lay_of<LAY>(a), x, cx.xp);
}

// Kernel body = fast path inline + out-of-line slow retry for the rare out-of-domain trajectory.  DIN = doubles loaded
// per trajectory.
//
// Work distribution: the batch is cut into TILES of 32 consecutive trajectories (one warp instruction moves one tile:
// contiguous 32 * 16n bytes); launched as ONE resident wave of G CTAs x W warps (runtime.cpp launch()), warp w of CTA b
// walks tiles b + G (w + W r), r = 0, 1, ... — every round spreads over all CTAs, so whatever the batch size the last,
// partial round leaves the same number of busy warps on every SM (a grid-stride loop over thread indices piles the
// remainder onto the first CTAs).  Trajectory indices are 32-bit (the host splits batches >= 2^31).
// Per launch, in this order:  griddepcontrol.launch_dependents (the next kernel of the stream may start its own
// prologue as soon as SM resources free up)  ->  sin/cos table staging by ONE cp.async.bulk (TMA) per CTA  ->  L2
// prefetch of the warp's first two tiles  ->  griddepcontrol.wait (everything above touches only constant data or
// prefetches into the coherent L2, so it overlaps the previous kernel's tail)  ->  the tile loop, which loads the
// thread's next Phase under the current trajectory's arithmetic into the idle one of two register buffers.
#define HB_PROCESS_(NAME, IDX, YBUF)                                                                       \
  {                                                                                                        \
    HbCtx cx;                                                                                              \
    cx.tab_s = tab_s;                                                                                      \
    cx.xp = xp;                                                                                            \
    hb_ctx_reset(cx);                                                                                      \
    HB_TRAJ_CALL_##NAME(IDX, YBUF)                                                                         \
    if (hb_retry(cx)) hb_slow_##NAME<S>(a, (long long)(IDX));                                              \
  }
#define HB_TRAJ_CALL_step_rk4(IDX, YBUF) hb_traj_step_rk4<S, true, LAY, unsigned, NST>(a, IDX, YBUF, w, cx);
#define HB_TRAJ_CALL_step_rkf45(IDX, YBUF) hb_traj_step_rkf45<S, true, LAY, unsigned>(a, IDX, YBUF, w, cx);
#define HB_TRAJ_CALL_evolve_rk4(IDX, YBUF) hb_traj_evolve_rk4<S, true, LAY, unsigned>(a, IDX, YBUF, w, cx);
#define HB_TRAJ_CALL_evolve_rkf45(IDX, YBUF) hb_traj_evolve_rkf45<S, true, LAY, unsigned>(a, IDX, YBUF, w, cx);
#define HB_TRAJ_CALL_ham_eqs(IDX, YBUF) hb_traj_ham_eqs<S, true, LAY, unsigned>(a, IDX, YBUF, w, cx);
#define HB_TRAJ_CALL_to_phase(IDX, YBUF) hb_traj_to_phase<S, true, LAY, unsigned>(a, IDX, YBUF, w, cx);
#define HB_TRAJ_CALL_from_phase(IDX, YBUF) hb_traj_from_phase<S, true, LAY, unsigned>(a, IDX, YBUF, w, cx);
#define HB_TRAJ_CALL_energies(IDX, YBUF) hb_traj_energies<S, true, LAY, unsigned>(a, IDX, YBUF, w, cx);
#define HB_TRAJ_CALL_upos(IDX, YBUF) hb_traj_upos<S, true, LAY, unsigned>(a, IDX, YBUF, w, cx);
// DIN / DOUT = doubles loaded / stored per trajectory.  big_tab: the statically allocated table of large systems.
#define HB_KERNEL_BODY(NAME, DIN_EXPR, DOUT_EXPR, STEPPING)                                                \
  template <class S>                                                                                       \
  __device__ __noinline__ void hb_slow_##NAME(const HbKArgs& a, long long i) {                             \
    constexpr int DIN = DIN_EXPR;                                                                          \
    double w[S::M], yin[DIN];                                                                              \
    S::inertia(a.prm, w);                                                                                  \
    hb_load<DIN, long long>(a.in, i, a.N, a.layout, yin);                                                  \
    HbCtx cx;                                                                                              \
    cx.tab_s = 0;                                                                                          \
    cx.xp = nullptr;                                                                                       \
    hb_ctx_reset(cx);                                                                                      \
    hb_traj_##NAME<S, false, -1, long long>(a, i, yin, w, cx);                                             \
  }                                                                                                        \
  template <class S, int LAY, int NST = 0>                                                                 \
  HB_DEV void hb_body_##NAME(const HbKArgs& a, HbTab* big_tab) {                                           \
    constexpr int DIN = DIN_EXPR, DOUT = DOUT_EXPR;                                                        \
    constexpr bool SMALL = S::N < HB_BIG_N;                                                                \
    /* cp.async staging only where a trajectory costs clearly more issue time than HBM time (Sys::HEAVY, stepping kernels) */ \
    constexpr bool ASYNC = SMALL && HB_ASYNC_STAGE && DIN % 2 == 0 && STEPPING && S::HEAVY;                \
    const unsigned N = (unsigned)a.N;                                                                      \
    const unsigned istride = gridDim.x * blockDim.x;   /* trajectories per round */                        \
    unsigned i = a.contiguous ? blockIdx.x * blockDim.x + threadIdx.x                                      \
                              : (blockIdx.x + gridDim.x * (threadIdx.x >> 5)) * 32u + (threadIdx.x & 31u); \
    const int lay = hb_lay_of<LAY>(a);                                                                     \
    /* carve the dynamic shared memory (layout: HB_DYN_DOUBLES) */                                         \
    HbTab* tab = SMALL ? reinterpret_cast<HbTab*>(hb_dsm) : big_tab;                                       \
    double* stage = hb_dsm + (S::TRIG ? HB_TAB_BYTES / 8 : 0);                                             \
    double* xp = SMALL ? stage + (ASYNC ? 2 * DIN * blockDim.x : 0) : hb_dsm + HB_DYN_DOUBLES(S::N, S::NE) * blockDim.x; \
    xp = (lay == 2 && DOUT % 2 == 0 && DOUT <= HB_WSTORE_MAXD) ? xp + (threadIdx.x & ~31u) * DOUT : nullptr; \
    HB_PDL_LAUNCH_DEPENDENTS();                                                                            \
    if constexpr (S::TRIG) hb_tab_issue(tab);                                                              \
    if (HB_PRE_L2 && !a.host_io) {                                                                         \
      if (i < N) hb_prefetch_l2<DIN, unsigned>(a.in, i, N, lay);                                           \
      if (i + istride < N) hb_prefetch_l2<DIN, unsigned>(a.in, i + istride, N, lay);                       \
    }                                                                                                      \
    if constexpr (S::TRIG) hb_tab_wait(tab);                                                               \
    HB_PDL_WAIT();                                                                                         \
    double w[S::M];                                                                                        \
    S::inertia(a.prm, w);                                                                                  \
    const unsigned tab_s = hb_smem_addr(tab);                                                              \
    bool more = i < N;                                                                                     \
    if (ASYNC && lay != 1 && !a.host_io) {   /* array of records in device memory: the next Phase lands in shared memory under this one's arithmetic (over PCIe cp.async runs at a quarter of the rate of plain loads: profiles/r2c) */ \
      const unsigned cstride = blockDim.x * 16u, sz = DIN * blockDim.x;   /* doubles per stage buffer */   \
      double* cur = stage + threadIdx.x * 2;                                                               \
      double* nxt = cur + sz;                                                                              \
      unsigned cur_s = hb_smem_addr(cur), nxt_s = cur_s + sz * 8u;                                         \
      if (more) hb_async_load<DIN>(cur, cur_s, cstride, a.in + (size_t)i * DIN);                           \
      while (more) {                                                                                       \
        double yin[DIN];                                                                                   \
        hb_async_read<DIN>(cur, cur_s, cstride, yin);                                                      \
        const unsigned inext = i + istride;                                                                \
        more = inext < N;                                                                                  \
        if (more) hb_async_load<DIN>(nxt, nxt_s, cstride, a.in + (size_t)inext * DIN);                     \
        HB_PROCESS_(NAME, i, yin)                                                                          \
        i = inext;                                                                                         \
        { double* tp = cur; cur = nxt; nxt = tp; const unsigned ts = cur_s; cur_s = nxt_s; nxt_s = ts; }   \
      }                                                                                                    \
    } else if constexpr (DIN <= 8 && HB_REG_PREFETCH) {   /* plain loads, small records: the next Phase is loaded into registers under this one's arithmetic (light kernels: the exposed L2 latency of a load at the top of every trajectory costs 8-30 %, profiles/r2n) */ \
      double yin[DIN];                                                                                     \
      if (more) hb_load<DIN, unsigned>(a.in, i, N, lay, yin);                                              \
      while (more) {                                                                                       \
        double ycur[DIN];                                                                                  \
        hb_copy<DIN>(yin, ycur);                                                                           \
        const unsigned inext = i + istride;                                                                \
        more = inext < N;                                                                                  \
        if (more) hb_load<DIN, unsigned>(a.in, inext, N, lay, yin);                                        \
        if (HB_PRE_L2 && !a.host_io) { if (inext + istride < N) hb_prefetch_l2<DIN, unsigned>(a.in, inext + istride, N, lay); } \
        HB_PROCESS_(NAME, i, ycur)                                                                         \
        i = inext;                                                                                         \
      }                                                                                                    \
    } else {                                                                                               \
      double yin[DIN];                                                                                     \
      while (more) {                                                                                       \
        hb_load<DIN, unsigned>(a.in, i, N, lay, yin);                                                      \
        if (HB_PRE_L2 && !a.host_io) { if (i + 2 * istride < N) hb_prefetch_l2<DIN, unsigned>(a.in, i + 2 * istride, N, lay); } \
        HB_PROCESS_(NAME, i, yin)                                                                          \
        i += istride;                                                                                      \
        more = i < N;                                                                                      \
      }                                                                                                    \
    }                                                                                                      \
  }
HB_KERNEL_BODY(step_rk4, 2 * S::N, 2 * S::N, true)
HB_KERNEL_BODY(step_rkf45, 2 * S::N, 2 * S::N, true)
HB_KERNEL_BODY(evolve_rk4, 2 * S::N, 2 * S::N, true)
HB_KERNEL_BODY(evolve_rkf45, 2 * S::N, 2 * S::N, true)
HB_KERNEL_BODY(ham_eqs, 2 * S::N, 2 * S::N, false)
HB_KERNEL_BODY(to_phase, 2 * S::N, 2 * S::N, false)
HB_KERNEL_BODY(from_phase, 2 * S::N, 2 * S::N, false)
HB_KERNEL_BODY(energies, 2 * S::N, 4, false)
HB_KERNEL_BODY(upos, S::N, S::M, false)

// Counter-based initial Phases; D = a.nsteps, lo = prm[0..D), hi = prm[D..2D)
HB_DEV double hb_splitmix_u01(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
HB_DEV void hb_body_init_random(const HbKArgs& a) {
  HB_PDL_LAUNCH_DEPENDENTS();
  HB_PDL_WAIT();   // every kernel of this library is launched with the PDL attribute
  const int D = a.nsteps;
  const long long total = a.N * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long i; int c;
    if (a.layout == 0) { i = e / D; c = (int)(e - i * D); } else { c = (int)(e / a.N); i = e - (long long)c * a.N; }
    const double u = hb_splitmix_u01(a.seed + (unsigned long long)D * (unsigned long long)(a.first + i) + (unsigned long long)c);
    // no FMA contraction: bit-identical to the host-side generator (lo + (hi - lo) * u with two roundings)
    a.out[e] = __dadd_rn(a.prm[c], __dmul_rn(__dsub_rn(a.prm[D + c], a.prm[c]), u));
  }
}

// Kernel ids: the order of every per-system kernel table.
#define HB_K_STEP_RK4 0
#define HB_K_STEP_RKF45 1
#define HB_K_EVOLVE_RK4 2
#define HB_K_EVOLVE_RKF45 3
#define HB_K_HAM_EQS 4
#define HB_K_TO_PHASE 5
#define HB_K_FROM_PHASE 6
#define HB_K_ENERGIES 7
#define HB_K_UPOS 8
#define HB_K_COUNT 9

#ifdef HB_MINB_RK4                              // tuning experiments: explicit register cap for the RK4 step kernel
#define HB_LB_RK4(SYS) __launch_bounds__(HB_MAXBLOCK_OF(SYS::N), HB_MINB_RK4)
#else   // default: no occupancy request — every explicit one measured slower at one step per launch (profiles/r1p)
#define HB_LB_RK4(SYS) __launch_bounds__(HB_MAXBLOCK_OF(SYS::N))
#endif

// Instantiates per-system __global__ kernels with C linkage names PFX_<kind>.  The sin/cos table image is declared here,
// once per kernel, and handed to the body (systems without sin/cos reserve nothing).
#define HB_TAB_DECL(SYS)                                                                          \
  __shared__ __align__(128) unsigned char hb_tab_raw[(SYS::TRIG && SYS::N >= HB_BIG_N) ? sizeof(HbTab) : 16]; \
  HbTab* hb_tab = reinterpret_cast<HbTab*>(hb_tab_raw)
// step kernels: one instance per layout class, picked once per launch (large systems: the step dwarfs the bookkeeping; one
// instance keeps compile time)
#define HB_DEFINE_KERNEL_L(SYS, PFX, KIND, LB, NST1)                                              \
  extern "C" __global__ void LB PFX##_##KIND(const __grid_constant__ HbKArgs a) {                 \
    HB_TAB_DECL(SYS);                                                                             \
    if constexpr (HB_LAYSPEC && SYS::N < HB_BIG_N) {                                              \
      if (a.layout != 0) hb_body_##KIND<SYS, -1>(a, hb_tab);                                      \
      else if (NST1 && a.nsteps == 1) hb_body_##KIND<SYS, 0, NST1>(a, hb_tab);                    \
      else hb_body_##KIND<SYS, 0>(a, hb_tab);                                                     \
    } else {                                                                                      \
      hb_body_##KIND<SYS, -1>(a, hb_tab);                                                         \
    } }
#define HB_DEFINE_KERNEL_step_rk4(SYS, PFX) HB_DEFINE_KERNEL_L(SYS, PFX, step_rk4, HB_LB_RK4(SYS), 1)
#define HB_DEFINE_KERNEL_X(SYS, PFX, KIND) extern "C" __global__ void __launch_bounds__(HB_MAXBLOCK_OF(SYS::N)) PFX##_##KIND(const __grid_constant__ HbKArgs a) { HB_TAB_DECL(SYS); hb_body_##KIND<SYS, -1>(a, hb_tab); }
#ifdef HB_MINB_RKF45   // tuning experiments: register cap for the adaptive kernels
#define HB_LB_RKF45(SYS) __launch_bounds__(HB_MAXBLOCK_OF(SYS::N), HB_MINB_RKF45)
#else
// The adaptive kernels keep six stage vectors live and reach 200+ registers (2 CTAs/SM: two warps per scheduler cannot
// cover the 8-clock FP64 latency).  Capping small systems at 128 registers (4 CTAs/SM) spills a few stage vectors to
// local memory and still wins: double pendulum +10 %, triple pendulum +11 % (profiles/r1z/ab_rkf45_regs.txt).
#ifndef HB_RKF45_CTAS
#define HB_RKF45_CTAS(NCOORD) 2     // CTAs of 256 threads per SM the adaptive kernels of small systems are compiled for (2: 128 registers)
#endif
#define HB_LB_RKF45(SYS) __launch_bounds__((SYS::N <= 3 ? 256 : HB_MAXBLOCK_OF(SYS::N)), (SYS::N <= 3 ? HB_RKF45_CTAS(SYS::N) : 0))
#endif
#define HB_DEFINE_KERNEL_Y(SYS, PFX, KIND) extern "C" __global__ void HB_LB_RKF45(SYS) PFX##_##KIND(const __grid_constant__ HbKArgs a) { HB_TAB_DECL(SYS); hb_body_##KIND<SYS, -1>(a, hb_tab); }
#define HB_DEFINE_KERNEL_step_rkf45(SYS, PFX) HB_DEFINE_KERNEL_L(SYS, PFX, step_rkf45, HB_LB_RKF45(SYS), 0)
#define HB_DEFINE_KERNEL_evolve_rk4(SYS, PFX) HB_DEFINE_KERNEL_X(SYS, PFX, evolve_rk4)
#define HB_DEFINE_KERNEL_evolve_rkf45(SYS, PFX) HB_DEFINE_KERNEL_Y(SYS, PFX, evolve_rkf45)
#define HB_DEFINE_KERNEL_ham_eqs(SYS, PFX) HB_DEFINE_KERNEL_X(SYS, PFX, ham_eqs)
#define HB_DEFINE_KERNEL_to_phase(SYS, PFX) HB_DEFINE_KERNEL_X(SYS, PFX, to_phase)
#define HB_DEFINE_KERNEL_from_phase(SYS, PFX) HB_DEFINE_KERNEL_X(SYS, PFX, from_phase)
#define HB_DEFINE_KERNEL_energies(SYS, PFX) HB_DEFINE_KERNEL_X(SYS, PFX, energies)
#define HB_DEFINE_KERNEL_upos(SYS, PFX) HB_DEFINE_KERNEL_X(SYS, PFX, upos)
#define HB_DEFINE_KERNELS(SYS, PFX)                                                                  \
  HB_DEFINE_KERNEL_step_rk4(SYS, PFX) HB_DEFINE_KERNEL_step_rkf45(SYS, PFX) HB_DEFINE_KERNEL_evolve_rk4(SYS, PFX)      \
  HB_DEFINE_KERNEL_evolve_rkf45(SYS, PFX) HB_DEFINE_KERNEL_ham_eqs(SYS, PFX) HB_DEFINE_KERNEL_to_phase(SYS, PFX)       \
  HB_DEFINE_KERNEL_from_phase(SYS, PFX) HB_DEFINE_KERNEL_energies(SYS, PFX) HB_DEFINE_KERNEL_upos(SYS, PFX)
