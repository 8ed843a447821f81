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
