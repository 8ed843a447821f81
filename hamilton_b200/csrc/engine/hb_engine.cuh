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
#ifndef HB_BLOCK
#define HB_BLOCK 128      // threads per CTA of every kernel (the host launches exactly this)
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
  int nsteps;             // steps per launch
  int layout;             // 0 = AOS y[i*D+c], 1 = SOA y[c*N+i]
  int s;                  // evolve: number of grid points
  int substeps;           // evolve/RK4: equal sub-steps per grid interval
  unsigned long long seed;
  long long first;
  double prm[HB_MAXP];    // runtime parameters (HB_OP_PARAM leaves); for init_random: lo[0..D), hi[0..D)
};

// ---------------------------------------------------------------------------- batch I/O ----
template <int D>
HB_DEV void hb_load(const double* __restrict__ base, long long i, long long N, int layout, double (&y)[D]) {
  if (layout == 0) {
    if constexpr (D % 2 == 0) {   // 16-byte vector loads: a Phase is 2n doubles, always even
      const double2* p = reinterpret_cast<const double2*>(base + i * D);
#pragma unroll
      for (int c = 0; c < D / 2; c++) { double2 v = p[c]; y[2 * c] = v.x; y[2 * c + 1] = v.y; }
    } else {
#pragma unroll
      for (int c = 0; c < D; c++) y[c] = base[i * D + c];
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; c++) y[c] = base[(long long)c * N + i];   // one coalesced stream per component
  }
}
template <int D>
HB_DEV void hb_store(double* __restrict__ base, long long i, long long N, int layout, const double (&y)[D]) {
  if (layout == 0) {
    if constexpr (D % 2 == 0) {
      double2* p = reinterpret_cast<double2*>(base + i * D);
#pragma unroll
      for (int c = 0; c < D / 2; c++) p[c] = make_double2(y[2 * c], y[2 * c + 1]);
    } else {
#pragma unroll
      for (int c = 0; c < D; c++) base[i * D + c] = y[c];
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; c++) base[(long long)c * N + i] = y[c];
  }
}
HB_DEV bool hb_finite(double x) { return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000; }
template <int D>
HB_DEV void hb_copy(const double* src, double (&dst)[D]) {
#pragma unroll
  for (int c = 0; c < D; c++) dst[c] = src[c];
}

// ------------------------------------------------------------------------ fp64 primitives --
// The FP64 pipe (64 FMA/clk/SM) and the issue slots are what bound this engine, so the two
// transcendental-class primitives every mechanical system leans on are hand-written.
//
// hb_sincos<FAST=true>: x = k*(pi/64) + r with k = rint(x*64/pi) obtained from the 1.5*2^52 magic
// constant (no F2I/I2F), two-term Cody-Waite reduction, |r| <= pi/128; (sin, cos)(k*pi/64) come from
// a 128-entry table staged in shared memory (one LDS.128), sin r and cos r - 1 from 3-term
// polynomials, combined by the angle-addition formulas: 16 DFMA-class instructions instead of the
// ~24 + ~28 immediate-materialising moves of libdevice's sincos, no quadrant logic, no branch.
// Arguments with |x| >= 1e5 (or non-finite) only set cx.oob; the caller then redoes that
// trajectory on the out-of-line slow path (FAST=false: libdevice sincos with Payne-Hanek reduction).
// Max abs error of the fast path ~2e-16 (checked against long double on the host).
struct HbCtx {
  unsigned tab_s;       // shared-window address of the staged hb_kSinCosTab (fast path only)
  unsigned oob;         // set when a fast-path primitive saw an argument outside its domain
};

static __device__ const double2 hb_kSinCosTab[128] = {   // {sin, cos}(k*pi/64), correctly rounded
    {0, 1}, {0.049067674327418015, 0.99879545620517241},
    {0.098017140329560604, 0.99518472667219693}, {0.14673047445536175, 0.98917650996478101},
    {0.19509032201612828, 0.98078528040323043}, {0.2429801799032639, 0.97003125319454397},
    {0.29028467725446239, 0.95694033573220882}, {0.33688985339222005, 0.94154406518302081},
    {0.38268343236508978, 0.92387953251128674}, {0.42755509343028208, 0.90398929312344334},
    {0.47139673682599764, 0.88192126434835505}, {0.51410274419322177, 0.85772861000027212},
    {0.55557023301960218, 0.83146961230254524}, {0.59569930449243336, 0.80320753148064494},
    {0.63439328416364549, 0.77301045336273699}, {0.67155895484701844, 0.74095112535495911},
    {0.70710678118654757, 0.70710678118654757}, {0.74095112535495911, 0.67155895484701844},
    {0.77301045336273699, 0.63439328416364549}, {0.80320753148064494, 0.59569930449243336},
    {0.83146961230254524, 0.55557023301960218}, {0.85772861000027212, 0.51410274419322177},
    {0.88192126434835505, 0.47139673682599764}, {0.90398929312344334, 0.42755509343028208},
    {0.92387953251128674, 0.38268343236508978}, {0.94154406518302081, 0.33688985339222005},
    {0.95694033573220882, 0.29028467725446239}, {0.97003125319454397, 0.2429801799032639},
    {0.98078528040323043, 0.19509032201612828}, {0.98917650996478101, 0.14673047445536175},
    {0.99518472667219693, 0.098017140329560604}, {0.99879545620517241, 0.049067674327418015},
    {1, 0}, {0.99879545620517241, -0.049067674327418015},
    {0.99518472667219693, -0.098017140329560604}, {0.98917650996478101, -0.14673047445536175},
    {0.98078528040323043, -0.19509032201612828}, {0.97003125319454397, -0.2429801799032639},
    {0.95694033573220882, -0.29028467725446239}, {0.94154406518302081, -0.33688985339222005},
    {0.92387953251128674, -0.38268343236508978}, {0.90398929312344334, -0.42755509343028208},
    {0.88192126434835505, -0.47139673682599764}, {0.85772861000027212, -0.51410274419322177},
    {0.83146961230254524, -0.55557023301960218}, {0.80320753148064494, -0.59569930449243336},
    {0.77301045336273699, -0.63439328416364549}, {0.74095112535495911, -0.67155895484701844},
    {0.70710678118654757, -0.70710678118654757}, {0.67155895484701844, -0.74095112535495911},
    {0.63439328416364549, -0.77301045336273699}, {0.59569930449243336, -0.80320753148064494},
    {0.55557023301960218, -0.83146961230254524}, {0.51410274419322177, -0.85772861000027212},
    {0.47139673682599764, -0.88192126434835505}, {0.42755509343028208, -0.90398929312344334},
    {0.38268343236508978, -0.92387953251128674}, {0.33688985339222005, -0.94154406518302081},
    {0.29028467725446239, -0.95694033573220882}, {0.2429801799032639, -0.97003125319454397},
    {0.19509032201612828, -0.98078528040323043}, {0.14673047445536175, -0.98917650996478101},
    {0.098017140329560604, -0.99518472667219693}, {0.049067674327418015, -0.99879545620517241},
    {0, -1}, {-0.049067674327418015, -0.99879545620517241},
    {-0.098017140329560604, -0.99518472667219693}, {-0.14673047445536175, -0.98917650996478101},
    {-0.19509032201612828, -0.98078528040323043}, {-0.2429801799032639, -0.97003125319454397},
    {-0.29028467725446239, -0.95694033573220882}, {-0.33688985339222005, -0.94154406518302081},
    {-0.38268343236508978, -0.92387953251128674}, {-0.42755509343028208, -0.90398929312344334},
    {-0.47139673682599764, -0.88192126434835505}, {-0.51410274419322177, -0.85772861000027212},
    {-0.55557023301960218, -0.83146961230254524}, {-0.59569930449243336, -0.80320753148064494},
    {-0.63439328416364549, -0.77301045336273699}, {-0.67155895484701844, -0.74095112535495911},
    {-0.70710678118654757, -0.70710678118654757}, {-0.74095112535495911, -0.67155895484701844},
    {-0.77301045336273699, -0.63439328416364549}, {-0.80320753148064494, -0.59569930449243336},
    {-0.83146961230254524, -0.55557023301960218}, {-0.85772861000027212, -0.51410274419322177},
    {-0.88192126434835505, -0.47139673682599764}, {-0.90398929312344334, -0.42755509343028208},
    {-0.92387953251128674, -0.38268343236508978}, {-0.94154406518302081, -0.33688985339222005},
    {-0.95694033573220882, -0.29028467725446239}, {-0.97003125319454397, -0.2429801799032639},
    {-0.98078528040323043, -0.19509032201612828}, {-0.98917650996478101, -0.14673047445536175},
    {-0.99518472667219693, -0.098017140329560604}, {-0.99879545620517241, -0.049067674327418015},
    {-1, 0}, {-0.99879545620517241, 0.049067674327418015},
    {-0.99518472667219693, 0.098017140329560604}, {-0.98917650996478101, 0.14673047445536175},
    {-0.98078528040323043, 0.19509032201612828}, {-0.97003125319454397, 0.2429801799032639},
    {-0.95694033573220882, 0.29028467725446239}, {-0.94154406518302081, 0.33688985339222005},
    {-0.92387953251128674, 0.38268343236508978}, {-0.90398929312344334, 0.42755509343028208},
    {-0.88192126434835505, 0.47139673682599764}, {-0.85772861000027212, 0.51410274419322177},
    {-0.83146961230254524, 0.55557023301960218}, {-0.80320753148064494, 0.59569930449243336},
    {-0.77301045336273699, 0.63439328416364549}, {-0.74095112535495911, 0.67155895484701844},
    {-0.70710678118654757, 0.70710678118654757}, {-0.67155895484701844, 0.74095112535495911},
    {-0.63439328416364549, 0.77301045336273699}, {-0.59569930449243336, 0.80320753148064494},
    {-0.55557023301960218, 0.83146961230254524}, {-0.51410274419322177, 0.85772861000027212},
    {-0.47139673682599764, 0.88192126434835505}, {-0.42755509343028208, 0.90398929312344334},
    {-0.38268343236508978, 0.92387953251128674}, {-0.33688985339222005, 0.94154406518302081},
    {-0.29028467725446239, 0.95694033573220882}, {-0.2429801799032639, 0.97003125319454397},
    {-0.19509032201612828, 0.98078528040323043}, {-0.14673047445536175, 0.98917650996478101},
    {-0.098017140329560604, 0.99518472667219693}, {-0.049067674327418015, 0.99879545620517241},
};
static __device__ __constant__ double hb_kSC[10] = {
    20.371832715762604,           // 0  64/pi
    6755399441055744.0,           // 1  1.5 * 2^52
    0.04908738521234052,          // 2  pi/64 hi
    1.9135106236677394e-18,       // 3  pi/64 lo
    -1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0,      // 4..6   sin r = r + r^3 (S1 + z (S2 + z S3))
    -0.5, 1.0 / 24.0, -1.0 / 720.0};             // 7..9   cos r - 1 = z (C1 + z (C2 + z C3))

HB_DEV void hb_tab_init(double2* tab) {   // kernels are always launched with HB_BLOCK threads (compile-time stride)
#pragma unroll
  for (int t = threadIdx.x; t < 128; t += HB_BLOCK) tab[t] = hb_kSinCosTab[t];
  __syncthreads();
}

template <bool FAST>
HB_DEV void hb_sincos(HbCtx& cx, double x, double* sp, double* cp) {
  if constexpr (!FAST) {
    sincos(x, sp, cp);
  } else {
    cx.oob |= (unsigned)((__double2hiint(x) & 0x7fffffff) >= 0x40F86A00);   // |x| >= 1e5, inf or nan (integer pipe)
    const double t = fma(x, hb_kSC[0], hb_kSC[1]);
    double2 sc;   // one LDS.128 with a 32-bit shared address (no generic->shared conversion in the loop)
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(sc.x), "=d"(sc.y) : "r"(cx.tab_s + ((__double2loint(t) & 127) << 4)));
    const double kf = t - hb_kSC[1];
    double r = fma(-kf, hb_kSC[2], x);
    r = fma(-kf, hb_kSC[3], r);
    const double z = r * r;
    double ps = fma(z, hb_kSC[6], hb_kSC[5]);
    double pc = fma(z, hb_kSC[9], hb_kSC[8]);
    ps = fma(z, ps, hb_kSC[4]);
    pc = fma(z, pc, hb_kSC[7]);
    const double sr = fma(z * r, ps, r);      // sin r
    const double cm = z * pc;                 // cos r - 1
    *sp = fma(sc.y, sr, fma(sc.x, cm, sc.x));    // sin(a + r) = sin a + (sin a (cos r - 1) + cos a sin r)
    *cp = fma(-sc.x, sr, fma(sc.y, cm, sc.y));   // cos(a + r) = cos a + (cos a (cos r - 1) - sin a sin r)
  }
}
template <bool FAST> HB_DEV double hb_sin(HbCtx& cx, double x) { double s, c; hb_sincos<FAST>(cx, x, &s, &c); return s; }
template <bool FAST> HB_DEV double hb_cos(HbCtx& cx, double x) { double s, c; hb_sincos<FAST>(cx, x, &s, &c); return c; }

// hb_rcp: 1/d for the mass-matrix pivots.  MUFU.RCP64H seed + two Newton steps (<= 1 ulp); no
// denormal/overflow slow path: a pivot that needs one is flagged HB_FLAG_NOT_SPD / NONFINITE anyway.
HB_DEV double hb_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
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

// Solve A x = b for the packed SPD mass matrix; A is destroyed.  Replaces the reference's explicit
// `inv jmj` (src/Numeric/Hamilton.hs:381); a non-positive pivot is the analogue of hmatrix's
// singular-matrix exception.  N = 1, 2: closed form with ONE reciprocal (shortest dependency chain);
// N >= 3: LDL^T — no square roots, N reciprocals — then forward/diagonal/backward substitution.
// pivot test on the integer pipe: true unless d is a positive normal number (NaN passes here and is
// caught by the non-finite check on the result)
HB_DEV bool hb_bad_pivot(double d) { return __double2hiint(d) < 0x00100000; }

template <int N>
HB_DEV void hb_spd_solve(double* A, const double* b, double* x, int& flag) {
  if constexpr (N == 1) {
    if (hb_bad_pivot(A[0])) flag |= HB_FLAG_NOT_SPD;
    x[0] = b[0] * hb_rcp(A[0]);
  } else if constexpr (N == 2) {
    const double det = fma(A[0], A[2], -A[1] * A[1]);
    if (hb_bad_pivot(A[0]) || hb_bad_pivot(det)) flag |= HB_FLAG_NOT_SPD;
    const double id = hb_rcp(det);
    const double n0 = fma(A[2], b[0], -A[1] * b[1]);
    const double n1 = fma(A[0], b[1], -A[1] * b[0]);
    x[0] = n0 * id;
    x[1] = n1 * id;
  } else {
    double invd[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
      double v[N];
      double d = A[hb_tri(j, j)];
#pragma unroll
      for (int k = 0; k < j; k++) {
        v[k] = A[hb_tri(j, k)] * A[hb_tri(k, k)];   // L_jk d_k   (A_kk holds d_k)
        d = fma(-A[hb_tri(j, k)], v[k], d);
      }
      if (hb_bad_pivot(d)) flag |= HB_FLAG_NOT_SPD;
      A[hb_tri(j, j)] = d;
      const double id = hb_rcp(d);
      invd[j] = id;
#pragma unroll
      for (int i = j + 1; i < N; i++) {
        double t = A[hb_tri(i, j)];
#pragma unroll
        for (int k = 0; k < j; k++) t = fma(-A[hb_tri(i, k)], v[k], t);
        A[hb_tri(i, j)] = t * id;
      }
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
      double t = b[j];
#pragma unroll
      for (int k = 0; k < j; k++) t = fma(-A[hb_tri(j, k)], x[k], t);
      x[j] = t;
    }
#pragma unroll
    for (int j = 0; j < N; j++) x[j] *= invd[j];
#pragma unroll
    for (int j = N - 1; j >= 0; j--) {
      double t = x[j];
#pragma unroll
      for (int k = j + 1; k < N; k++) t = fma(-A[hb_tri(k, j)], x[k], t);
      x[j] = t;
    }
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
  double Jv[NJ > 0 ? NJ : 1], Hv[NH > 0 ? NH : 1], gU[N];
  double qq[N];
#pragma unroll
  for (int j = 0; j < N; j++) qq[j] = q[j];
  S::template derivs<FAST>(cx, prm, qq, Jv, Hv, gU);
  double wJ[NJ > 0 ? NJ : 1];
  hb_weigh<S>(w, Jv, wJ);
  double A[N * (N + 1) / 2];
  hb_mass<S>(wJ, Jv, A);
  hb_spd_solve<N>(A, p, dq, flag);
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
  if constexpr (WITH_U) S::template jac_pot<FAST>(cx, prm, qq, Jv, U); else S::template jac<FAST>(cx, prm, qq, Jv);
  hb_weigh<S>(w, Jv, wJ);
  double A[N * (N + 1) / 2];
  hb_mass<S>(wJ, Jv, A);
  hb_spd_solve<N>(A, p, v, flag);
}

// ------------------------------------------------------------------------ integrators ------
// Classical RK4 — the unit of BASELINE.json's metric (4 hamEqs evaluations per step).
template <class S, bool FAST>
HB_DEV void hb_rk4_step(HbCtx& cx, const double* prm, const double* w, double (&y)[2 * S::N], double dt, double h6, int& flag) {
  constexpr int D = 2 * S::N;
  double k[D], yt[D], acc[D];
  const double hh = 0.5 * dt;
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
template <class S, bool FAST>
HB_DEV void hb_rkf45_apply(HbCtx& cx, const double* prm, const double* w, double h, double (&y)[2 * S::N],
                           const double (&k1)[2 * S::N], double (&yerr)[2 * S::N],
                           double (&dydt_out)[2 * S::N], int& flag) {
  constexpr int D = 2 * S::N;
  typedef HbRkf45 T;
  double k2[D], k3[D], k4[D], k5[D], k6[D], yt[D];
#pragma unroll
  for (int c = 0; c < D; c++) yt[c] = y[c] + T::ah0 * h * k1[c];
  hb_rhs<S, FAST>(cx, prm, w, yt, k2, flag);
#pragma unroll
  for (int c = 0; c < D; c++) yt[c] = y[c] + h * (T::b30 * k1[c] + T::b31 * k2[c]);
  hb_rhs<S, FAST>(cx, prm, w, yt, k3, flag);
#pragma unroll
  for (int c = 0; c < D; c++) yt[c] = y[c] + h * (T::b40 * k1[c] + T::b41 * k2[c] + T::b42 * k3[c]);
  hb_rhs<S, FAST>(cx, prm, w, yt, k4, flag);
#pragma unroll
  for (int c = 0; c < D; c++) yt[c] = y[c] + h * (T::b50 * k1[c] + T::b51 * k2[c] + T::b52 * k3[c] + T::b53 * k4[c]);
  hb_rhs<S, FAST>(cx, prm, w, yt, k5, flag);
#pragma unroll
  for (int c = 0; c < D; c++)
    yt[c] = y[c] + h * (T::b60 * k1[c] + T::b61 * k2[c] + T::b62 * k3[c] + T::b63 * k4[c] + T::b64 * k5[c]);
  hb_rhs<S, FAST>(cx, prm, w, yt, k6, flag);
#pragma unroll
  for (int c = 0; c < D; c++) {
    const double di = T::c1 * k1[c] + T::c3 * k3[c] + T::c4 * k4[c] + T::c5 * k5[c] + T::c6 * k6[c];
    y[c] += h * di;
    yerr[c] = h * (T::e1 * k1[c] + T::e3 * k3[c] + T::e4 * k4[c] + T::e5 * k5[c] + T::e6 * k6[c]);
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
  int guard = 0;
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
      // std_control_hadjust, ord = 5
      const double h_old = h0;
      double rmax = 2.2250738585072014e-308;   // DBL_MIN
#pragma unroll
      for (int c = 0; c < D; c++) {
        const double D0 = T::eps * (fabs(y[c]) + fabs(h_old * dout[c])) + T::eps;
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
// Each hb_traj_* processes trajectory i completely (load -> compute -> store).  FAST=true is inlined
// into the kernel; if any fast primitive left its domain (cx.oob) nothing is stored and the kernel
// re-runs that one trajectory through the out-of-line FAST=false instance (HB_KERNEL_BODY below).
template <int D>
HB_DEV void hb_finish(const HbKArgs& a, long long i, const double (&y)[D], int flag) {
  bool ok = true;
#pragma unroll
  for (int c = 0; c < D; c++) ok = ok && hb_finite(y[c]);
  if (!ok) flag |= HB_FLAG_NONFINITE;
  if (flag && a.flags) a.flags[i] |= flag;
}

// stepHam iterated with the fixed RK4 stepper
template <class S, bool FAST>
HB_DEV void hb_traj_step_rk4(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N;
  double y[D];
  hb_copy<D>(yin, y);
  int flag = 0;
  for (int s = 0; s < a.nsteps; s++) hb_rk4_step<S, FAST>(cx, a.prm, w, y, a.dt, a.dt6, flag);
  if (FAST && cx.oob) return;
  hb_store<D>(a.out, i, a.N, a.layout, y);
  hb_finish<D>(a, i, y, flag);
}
// stepHam iterated with reference semantics: each step is a fresh adaptive solve over (0, dt)
template <class S, bool FAST>
HB_DEV void hb_traj_step_rkf45(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N;
  double y[D];
  hb_copy<D>(yin, y);
  int flag = 0;
  for (int s = 0; s < a.nsteps; s++) {
    HbEvolve<D> e;
    e.h = a.dt / 100;   // hi = (t1 - t0)/100, src/Numeric/Hamilton.hs:447
    e.primed = false;
    double t = 0.0;
    hb_rkf45_to<S, FAST>(cx, a.prm, w, y, t, a.dt, e, flag);
  }
  if (FAST && cx.oob) return;
  hb_store<D>(a.out, i, a.N, a.layout, y);
  hb_finish<D>(a, i, y, flag);
}
// evolveHam over a shared time grid; out[k] = batch at ts[k]
template <class S, bool FAST, bool ADAPTIVE>
HB_DEV void hb_traj_evolve(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N;
  double y[D];
  hb_copy<D>(yin, y);
  if (!FAST || !cx.oob) hb_store<D>(a.out, i, a.N, a.layout, y);   // row 0 is the initial state
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
      for (int s = 0; s < a.substeps; s++) hb_rk4_step<S, FAST>(cx, a.prm, w, y, h, h6, flag);
      t = tk;
    }
    if (FAST && cx.oob) return;   // rows written so far are rewritten by the slow retry (out never aliases in)
    hb_store<D>(a.out + (long long)k * a.N * D, i, a.N, a.layout, y);
  }
  hb_finish<D>(a, i, y, flag);
}
template <class S, bool FAST>
HB_DEV void hb_traj_evolve_rk4(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) { hb_traj_evolve<S, FAST, false>(a, i, yin, w, cx); }
template <class S, bool FAST>
HB_DEV void hb_traj_evolve_rkf45(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) { hb_traj_evolve<S, FAST, true>(a, i, yin, w, cx); }

template <class S, bool FAST>
HB_DEV void hb_traj_ham_eqs(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N;
  double y[D], dy[D];
  hb_copy<D>(yin, y);
  int flag = 0;
  hb_rhs<S, FAST>(cx, a.prm, w, y, dy, flag);
  if (FAST && cx.oob) return;
  hb_store<D>(a.out, i, a.N, a.layout, dy);
  hb_finish<D>(a, i, dy, flag);
}
template <class S, bool FAST>
HB_DEV void hb_traj_to_phase(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) {   // Config [q, v] -> Phase [q, p]
  constexpr int D = 2 * S::N, N = S::N;
  double c[D], y[D];
  hb_copy<D>(yin, c);
#pragma unroll
  for (int j = 0; j < N; j++) y[j] = c[j];
  hb_momenta<S, FAST>(cx, a.prm, w, c, c + N, y + N);
  if (FAST && cx.oob) return;
  hb_store<D>(a.out, i, a.N, a.layout, y);
}
template <class S, bool FAST>
HB_DEV void hb_traj_from_phase(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) {   // Phase [q, p] -> Config [q, v]
  constexpr int D = 2 * S::N, N = S::N;
  double y[D], c[D];
  hb_copy<D>(yin, y);
  int flag = 0;
  double U;
#pragma unroll
  for (int j = 0; j < N; j++) c[j] = y[j];
  hb_velocities<S, FAST, false>(cx, a.prm, w, y, y + N, c + N, U, flag);
  if (FAST && cx.oob) return;
  hb_store<D>(a.out, i, a.N, a.layout, c);
  hb_finish<D>(a, i, c, flag);
}
// out4[i] = (keP, pe, hamiltonian, lagrangian)
template <class S, bool FAST>
HB_DEV void hb_traj_energies(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) {
  constexpr int D = 2 * S::N, N = S::N;
  double y[D], v[N];
  hb_copy<D>(yin, y);
  int flag = 0;
  double U;
  hb_velocities<S, FAST, true>(cx, a.prm, w, y, y + N, v, U, flag);
  if (FAST && cx.oob) return;
  double T = 0.0;
#pragma unroll
  for (int j = 0; j < N; j++) T = fma(v[j], y[N + j], T);
  T *= 0.5;   // (vs <.> ps) / 2, src/Numeric/Hamilton.hs:349
  double o[4] = {T, U, T + U, T - U};
  hb_store<4>(a.out, i, a.N, 0, o);
  hb_finish<4>(a, i, o, flag);
}
template <class S, bool FAST>
HB_DEV void hb_traj_upos(const HbKArgs& a, long long i, const double* yin, const double* w, HbCtx& cx) {   // underlyingPos
  constexpr int N = S::N, M = S::M;
  double q[N], x[M];
  (void)w;
  hb_copy<N>(yin, q);
  S::template pos<FAST>(cx, a.prm, q, x);
  if (FAST && cx.oob) return;
  hb_store<M>(a.out, i, a.N, a.layout, x);
}

// Kernel body = fast path inline + out-of-line slow retry for the rare out-of-domain trajectory.
// DIN = doubles loaded per trajectory.  The trajectory's input is requested from HBM BEFORE the
// shared-memory table is staged, so the table's LDG->STS->barrier chain hides under the DRAM latency;
// when a thread owns several trajectories (grid-stride), the next input is prefetched under the current compute.
#define HB_KERNEL_BODY(NAME, DIN_EXPR)                                                                     \
  template <class S>                                                                                       \
  __device__ __noinline__ void hb_slow_##NAME(const HbKArgs& a, long long i) {                             \
    constexpr int DIN = DIN_EXPR;                                                                          \
    double w[S::M], yin[DIN];                                                                              \
    S::inertia(a.prm, w);                                                                                  \
    hb_load<DIN>(a.in, i, a.N, a.layout, yin);                                                             \
    HbCtx cx;                                                                                              \
    cx.tab_s = 0;                                                                                          \
    cx.oob = 0;                                                                                            \
    hb_traj_##NAME<S, false>(a, i, yin, w, cx);                                                            \
  }                                                                                                        \
  template <class S>                                                                                       \
  HB_DEV void hb_body_##NAME(const HbKArgs& a) {                                                           \
    constexpr int DIN = DIN_EXPR;                                                                          \
    __shared__ double2 tab[128];                                                                           \
    const long long stride = (long long)gridDim.x * blockDim.x;                                            \
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;                                        \
    double yin[DIN];                                                                                       \
    if (i < a.N) hb_load<DIN>(a.in, i, a.N, a.layout, yin);                                                \
    if constexpr (S::TRIG) hb_tab_init(tab);                                                               \
    double w[S::M];                                                                                        \
    S::inertia(a.prm, w);                                                                                  \
    const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);                                        \
    while (i < a.N) {                                                                                      \
      HbCtx cx;                                                                                            \
      cx.tab_s = tab_s;                                                                                    \
      cx.oob = 0;                                                                                          \
      if constexpr (DIN <= 8) {   /* small state: prefetch the thread's next trajectory under this one's compute */ \
        double ycur[DIN];                                                                                  \
        hb_copy<DIN>(yin, ycur);                                                                           \
        if (i + stride < a.N) hb_load<DIN>(a.in, i + stride, a.N, a.layout, yin);                          \
        hb_traj_##NAME<S, true>(a, i, ycur, w, cx);                                                        \
      } else {                                                                                             \
        hb_traj_##NAME<S, true>(a, i, yin, w, cx);                                                         \
      }                                                                                                    \
      if (cx.oob) hb_slow_##NAME<S>(a, i);                                                                 \
      i += stride;                                                                                         \
      if constexpr (DIN > 8) { if (i < a.N) hb_load<DIN>(a.in, i, a.N, a.layout, yin); }                   \
    }                                                                                                      \
  }
HB_KERNEL_BODY(step_rk4, 2 * S::N)
HB_KERNEL_BODY(step_rkf45, 2 * S::N)
HB_KERNEL_BODY(evolve_rk4, 2 * S::N)
HB_KERNEL_BODY(evolve_rkf45, 2 * S::N)
HB_KERNEL_BODY(ham_eqs, 2 * S::N)
HB_KERNEL_BODY(to_phase, 2 * S::N)
HB_KERNEL_BODY(from_phase, 2 * S::N)
HB_KERNEL_BODY(energies, 2 * S::N)
HB_KERNEL_BODY(upos, S::N)

// Counter-based initial Phases; D = a.nsteps, lo = prm[0..D), hi = prm[D..2D)
HB_DEV double hb_splitmix_u01(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
HB_DEV void hb_body_init_random(const HbKArgs& a) {
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

#ifdef HB_MINB_RK4   // optional register cap for the RK4 step kernel: 65536 / (HB_BLOCK * HB_MINB_RK4) registers/thread
#define HB_LB_RK4 __launch_bounds__(HB_BLOCK, HB_MINB_RK4)
#else
#define HB_LB_RK4 __launch_bounds__(HB_BLOCK)
#endif

// Instantiates the per-system __global__ kernels with C linkage names PFX_<kind>.
#define HB_DEFINE_KERNELS(SYS, PFX)                                                                         \
  extern "C" __global__ void HB_LB_RK4 PFX##_step_rk4(const __grid_constant__ HbKArgs a) { hb_body_step_rk4<SYS>(a); }      \
  extern "C" __global__ void __launch_bounds__(HB_BLOCK) PFX##_step_rkf45(const __grid_constant__ HbKArgs a) { hb_body_step_rkf45<SYS>(a); }  \
  extern "C" __global__ void __launch_bounds__(HB_BLOCK) PFX##_evolve_rk4(const __grid_constant__ HbKArgs a) { hb_body_evolve_rk4<SYS>(a); } \
  extern "C" __global__ void __launch_bounds__(HB_BLOCK) PFX##_evolve_rkf45(const __grid_constant__ HbKArgs a) { hb_body_evolve_rkf45<SYS>(a); } \
  extern "C" __global__ void __launch_bounds__(HB_BLOCK) PFX##_ham_eqs(const __grid_constant__ HbKArgs a) { hb_body_ham_eqs<SYS>(a); }        \
  extern "C" __global__ void __launch_bounds__(HB_BLOCK) PFX##_to_phase(const __grid_constant__ HbKArgs a) { hb_body_to_phase<SYS>(a); }      \
  extern "C" __global__ void __launch_bounds__(HB_BLOCK) PFX##_from_phase(const __grid_constant__ HbKArgs a) { hb_body_from_phase<SYS>(a); }  \
  extern "C" __global__ void __launch_bounds__(HB_BLOCK) PFX##_energies(const __grid_constant__ HbKArgs a) { hb_body_energies<SYS>(a); }      \
  extern "C" __global__ void __launch_bounds__(HB_BLOCK) PFX##_upos(const __grid_constant__ HbKArgs a) { hb_body_upos<SYS>(a); }
