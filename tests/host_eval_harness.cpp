// Compiles a generated `Sys` struct (hb_system_source) as HOST code so the system compiler's output
// (symbolic derivatives, sparsity tables) can be checked against the oracle without a GPU.
// Test infrastructure: the product never evaluates systems on the CPU.
#include <cmath>
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
struct HbCtx { int oob; };
template <bool FAST> static inline void hb_sincos(HbCtx&, double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
template <bool FAST> static inline double hb_sin(HbCtx&, double x) { return std::sin(x); }
template <bool FAST> static inline double hb_cos(HbCtx&, double x) { return std::cos(x); }
template <bool FAST> static inline double hb_recip(HbCtx&, double x) { return 1.0 / x; }
template <bool FAST> static inline double hb_exp(HbCtx&, double x) { return std::exp(x); }
using std::exp; using std::log; using std::sqrt; using std::pow; using std::fabs; using std::tan; using std::atan2;
using std::asin; using std::acos; using std::atan; using std::sinh; using std::cosh; using std::tanh;
using std::asinh; using std::acosh; using std::atanh;
#include SYS_SOURCE
typedef SYS_NAME S;
extern "C" void dims(int* o) { o[0] = S::M; o[1] = S::N; o[2] = S::NJ; o[3] = S::NH; }
// J: m x n dense; H: n x m x n dense (slice j, row i, col k); gU: n; x: m; U; w: m
extern "C" void eval(const double* prm, const double* q, double* J, double* H, double* gU, double* x, double* U, double* w) {
  constexpr int M = S::M, N = S::N;
  double Jv[S::NJ + 1], Hv[S::NH + 1], Jv2[S::NJ + 1], Jv3[S::NJ + 1];
  for (int i = 0; i < M * N; i++) J[i] = 0;
  for (int i = 0; i < N * M * N; i++) H[i] = 0;
  HbCtx cx{0};
  S::derivs<true>(cx, prm, q, Jv, Hv, gU);
  S::jac<false>(cx, prm, q, Jv2);
  S::jac_pot<true>(cx, prm, q, Jv3, *U);
  S::pos<true>(cx, prm, q, x);
  S::inertia(prm, w);
  for (int e = 0; e < S::NJ; e++) {
    J[S::jrow(e) * N + S::jcol(e)] = Jv[e];
    if (Jv2[e] != Jv[e] || Jv3[e] != Jv[e]) J[S::jrow(e) * N + S::jcol(e)] = NAN;   // the three emitters must agree
    if (S::jidx(S::jrow(e), S::jcol(e)) != e) J[0] = NAN;
  }
  for (int e = 0; e < S::NH; e++) {
    H[(S::hj(e) * M + S::hrow(e)) * N + S::hk(e)] = Hv[e];
    H[(S::hk(e) * M + S::hrow(e)) * N + S::hj(e)] = Hv[e];
  }
}

// Symbolic hamEqs (hpre -> dense SPD solve -> hpost), momenta via smass, and U via smass_pot.
// Returns 0 when the system was compiled with the direct contraction only (SYMH == false).
template <class T, bool ON = T::SYMH> struct SymEval {
  static int run(const double*, const double*, const double*, double*, double*, double*, double*) { return 0; }
};
template <class T> struct SymEval<T, true> {
  static int run(const double* prm, const double* q, const double* p, double* dq, double* dp, double* Aout, double* U) {
    constexpr int N = T::N, NT = N * (N + 1) / 2;
    double A[NT], A2[NT], A3[NT], E[T::NE + 1];
    HbCtx cx{0};
    T::template hpre<true>(cx, prm, q, A, E);
    T::template smass<true>(cx, prm, q, A2);
    T::template smass_pot<true>(cx, prm, q, A3, *U);
    double Md[N][N], b[N];
    for (int j = 0; j < N; j++)
      for (int k = 0; k <= j; k++) {
        const int t = j * (j + 1) / 2 + k;
        Md[j][k] = Md[k][j] = (A2[t] == A[t] && A3[t] == A[t]) ? A[t] : NAN;   // the three emitters must agree
        Aout[j * N + k] = Aout[k * N + j] = Md[j][k];
      }
    for (int j = 0; j < N; j++) b[j] = p[j];
    for (int c = 0; c < N; c++) {   // Gaussian elimination with partial pivoting (test harness only)
      int piv = c;
      for (int r = c + 1; r < N; r++) if (std::fabs(Md[r][c]) > std::fabs(Md[piv][c])) piv = r;
      for (int k = 0; k < N; k++) { double t = Md[c][k]; Md[c][k] = Md[piv][k]; Md[piv][k] = t; }
      { double t = b[c]; b[c] = b[piv]; b[piv] = t; }
      for (int r = c + 1; r < N; r++) {
        const double f = Md[r][c] / Md[c][c];
        for (int k = c; k < N; k++) Md[r][k] -= f * Md[c][k];
        b[r] -= f * b[c];
      }
    }
    for (int r = N - 1; r >= 0; r--) {
      double t = b[r];
      for (int k = r + 1; k < N; k++) t -= Md[r][k] * dq[k];
      dq[r] = t / Md[r][r];
    }
    T::template hpost<true>(cx, prm, q, E, dq, dp);
    return 1;
  }
};
extern "C" int eval_symham(const double* prm, const double* q, const double* p, double* dq, double* dp, double* A, double* U) {
  return SymEval<S>::run(prm, q, p, dq, dp, A, U);
}
