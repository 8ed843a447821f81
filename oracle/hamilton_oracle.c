/*
 * hamilton_oracle.c — CPU restatement of mstksg/hamilton's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (hamilton_b200/, include/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors or tests (test/Spec.hs:1-2 is a stub),
 * and neither GHC nor GSL exist in this image, so the Haskell binary cannot be run.  This file
 * restates the reference algorithm literally (plain dense arithmetic, no cleverness) and is
 * cross-checked by an independent sympy/scipy derivation (oracle/crosscheck.py) and physics
 * invariants; tests/golden/ holds vectors produced by THIS file, not by the reference.
 *
 * What is restated, with the reference lines followed (paths relative to /root/reference):
 *   - ad's jacobianT / hessianF / grad as used by mkSystem (src/Numeric/Hamilton.hs:217-225):
 *     dense second-order forward-mode jets (value, gradient, full Hessian) — exact derivatives,
 *     like `ad`, up to rounding.
 *   - mkSystem' = mkSystem m f (u . f) (:254).
 *   - momenta (:267), velocities (:321-324, explicit inverse like hmatrix `inv`), keC/keP/pe/
 *     lagrangian/hamiltonian (:288-361), hamEqs (:375-387, same association order of the #> chain).
 *   - evolveHam (:443-462): GSL rkf45 + standard controller + gsl_odeiv2_evolve_apply as driven by
 *     hmatrix-gsl's `odeSolveV RKf45 hi eps eps` (third-party, not vendored; semantics restated
 *     from GSL 2.x ode-initval2/{rkf45.c,cstd.c,evolve.c} and hmatrix-gsl 0.19 gsl-ode.c);
 *     stepHam (:400-402) = evolveHam over (0, r), row 1.
 *   - classical fixed-step RK4 (the unit of BASELINE.json's metric; not what the reference runs).
 *   - the example systems of app/Examples.hs:61-183 and helpers :601-627, as native C on jets,
 *     plus a tape interpreter (include/hamilton_b200.h hb_op) for user-defined systems.
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "../include/hamilton_b200.h"

#define NMAX HB_MAX_N
#define MMAX HB_MAX_M

/* ------------------------------------------------------------------ jets (replaces `ad`) --- */
typedef struct { double v; double g[NMAX]; double h[NMAX][NMAX]; } jet;

static void j_const(int n, jet* r, double c) {
  r->v = c;
  for (int j = 0; j < n; j++) { r->g[j] = 0; for (int k = 0; k < n; k++) r->h[j][k] = 0; }
}
static void j_var(int n, jet* r, double x, int idx) { j_const(n, r, x); r->g[idx] = 1.0; }
static void j_add(int n, jet* r, const jet* a, const jet* b) {
  r->v = a->v + b->v;
  for (int j = 0; j < n; j++) { r->g[j] = a->g[j] + b->g[j]; for (int k = 0; k < n; k++) r->h[j][k] = a->h[j][k] + b->h[j][k]; }
}
static void j_sub(int n, jet* r, const jet* a, const jet* b) {
  r->v = a->v - b->v;
  for (int j = 0; j < n; j++) { r->g[j] = a->g[j] - b->g[j]; for (int k = 0; k < n; k++) r->h[j][k] = a->h[j][k] - b->h[j][k]; }
}
static void j_mul(int n, jet* r, const jet* a, const jet* b) {
  jet t;
  t.v = a->v * b->v;
  for (int j = 0; j < n; j++) {
    t.g[j] = a->v * b->g[j] + b->v * a->g[j];
    for (int k = 0; k < n; k++)
      t.h[j][k] = a->v * b->h[j][k] + b->v * a->h[j][k] + a->g[j] * b->g[k] + a->g[k] * b->g[j];
  }
  *r = t;
}
/* r = phi(a) given phi, phi', phi'' at a->v */
static void j_chain(int n, jet* r, const jet* a, double f0, double f1, double f2) {
  jet t;
  t.v = f0;
  for (int j = 0; j < n; j++) {
    t.g[j] = f1 * a->g[j];
    for (int k = 0; k < n; k++) t.h[j][k] = f1 * a->h[j][k] + f2 * a->g[j] * a->g[k];
  }
  *r = t;
}
static void j_scale(int n, jet* r, const jet* a, double c) { j_chain(n, r, a, c * a->v, c, 0.0); }
static void j_addc(int n, jet* r, const jet* a, double c) { jet t = *a; t.v += c; (void)n; *r = t; }
static void j_neg(int n, jet* r, const jet* a) { j_scale(n, r, a, -1.0); }
static void j_recip(int n, jet* r, const jet* a) {
  double x = a->v; j_chain(n, r, a, 1.0 / x, -1.0 / (x * x), 2.0 / (x * x * x));
}
static void j_div(int n, jet* r, const jet* a, const jet* b) { jet t; j_recip(n, &t, b); j_mul(n, r, a, &t); }
static void j_sin(int n, jet* r, const jet* a) { double s = sin(a->v), c = cos(a->v); j_chain(n, r, a, s, c, -s); }
static void j_cos(int n, jet* r, const jet* a) { double s = sin(a->v), c = cos(a->v); j_chain(n, r, a, c, -s, -c); }
static void j_exp(int n, jet* r, const jet* a) { double e = exp(a->v); j_chain(n, r, a, e, e, e); }
static void j_log(int n, jet* r, const jet* a) { double x = a->v; j_chain(n, r, a, log(x), 1.0 / x, -1.0 / (x * x)); }
static void j_sqrt(int n, jet* r, const jet* a) { double s = sqrt(a->v); j_chain(n, r, a, s, 0.5 / s, -0.25 / (s * a->v)); }
/* a ** c with a literal exponent: ad's (**) differentiates only the base when the exponent is a
 * known constant, so negative bases with integral exponents work (spring: x ** 2, app/Examples.hs:154) */
static void j_powc(int n, jet* r, const jet* a, double c) {
  double x = a->v;
  j_chain(n, r, a, pow(x, c), c * pow(x, c - 1.0), c * (c - 1.0) * pow(x, c - 2.0));
}
/* Num (^): exponentiation by repeated multiplication (bezierCurve, app/Examples.hs:618) */
static void j_powi(int n, jet* r, const jet* a, int k) {
  if (k == 0) { j_const(n, r, 1.0); return; }
  int neg = k < 0; if (neg) k = -k;
  jet acc; int have = 0; jet base = *a;
  while (k) {
    if (k & 1) { if (have) j_mul(n, &acc, &acc, &base); else { acc = base; have = 1; } }
    k >>= 1; if (k) j_mul(n, &base, &base, &base);
  }
  if (neg) j_recip(n, r, &acc); else *r = acc;
}
static void j_pow(int n, jet* r, const jet* a, const jet* b) { /* general a ** b = exp(b log a) */
  jet l; j_log(n, &l, a); j_mul(n, &l, &l, b); j_exp(n, r, &l); r->v = pow(a->v, b->v);
}
static void j_unary(int n, jet* r, const jet* a, int op) {
  double x = a->v, t, u;
  switch (op) {
    case HB_OP_NEG: j_neg(n, r, a); break;
    case HB_OP_RECIP: j_recip(n, r, a); break;
    case HB_OP_ABS: t = (x > 0) - (x < 0); j_chain(n, r, a, fabs(x), t, 0.0); break;
    case HB_OP_SIGNUM: j_const(n, r, (double)((x > 0) - (x < 0))); break;
    case HB_OP_SQRT: j_sqrt(n, r, a); break;
    case HB_OP_EXP: j_exp(n, r, a); break;
    case HB_OP_LOG: j_log(n, r, a); break;
    case HB_OP_SIN: j_sin(n, r, a); break;
    case HB_OP_COS: j_cos(n, r, a); break;
    case HB_OP_TAN: t = tan(x); u = 1.0 + t * t; j_chain(n, r, a, t, u, 2.0 * t * u); break;
    case HB_OP_ASIN: u = 1.0 - x * x; j_chain(n, r, a, asin(x), 1.0 / sqrt(u), x / (u * sqrt(u))); break;
    case HB_OP_ACOS: u = 1.0 - x * x; j_chain(n, r, a, acos(x), -1.0 / sqrt(u), -x / (u * sqrt(u))); break;
    case HB_OP_ATAN: u = 1.0 + x * x; j_chain(n, r, a, atan(x), 1.0 / u, -2.0 * x / (u * u)); break;
    case HB_OP_SINH: j_chain(n, r, a, sinh(x), cosh(x), sinh(x)); break;
    case HB_OP_COSH: j_chain(n, r, a, cosh(x), sinh(x), cosh(x)); break;
    case HB_OP_TANH: t = tanh(x); u = 1.0 - t * t; j_chain(n, r, a, t, u, -2.0 * t * u); break;
    case HB_OP_ASINH: u = x * x + 1.0; j_chain(n, r, a, asinh(x), 1.0 / sqrt(u), -x / (u * sqrt(u))); break;
    case HB_OP_ACOSH: u = x * x - 1.0; j_chain(n, r, a, acosh(x), 1.0 / sqrt(u), -x / (u * sqrt(u))); break;
    case HB_OP_ATANH: u = 1.0 - x * x; j_chain(n, r, a, atanh(x), 1.0 / u, 2.0 * x / (u * u)); break;
    default: j_const(n, r, NAN);
  }
}

/* ---------------------------------------------------------------------------- systems ------ */
typedef struct ho_system {
  int m, n, builtin;            /* builtin = hb_builtin id, or -1 for tapes */
  double w[MMAX];               /* _sysInertia */
  int np; double prm[HB_MAX_PARAMS];
  int u_on_cart;
  /* tape systems */
  int nf, nu, uout; hb_op* f; hb_op* u; int fout[MMAX];
} ho_system;

static int eval_tape(int n, const hb_op* ops, int nops, const jet* in, int nin, const double* prm, int np, jet* nodes) {
  for (int k = 0; k < nops; k++) {
    const hb_op* o = &ops[k];
    int bin = (o->op >= HB_OP_ADD && o->op <= HB_OP_DIV) || o->op == HB_OP_POW || o->op == HB_OP_ATAN2;
    if (o->op >= HB_OP_ADD && (o->a < 0 || o->a >= k)) return -1;
    if (bin && (o->b < 0 || o->b >= k)) return -1;
    switch (o->op) {
      case HB_OP_INPUT: if (o->a < 0 || o->a >= nin) return -1; nodes[k] = in[o->a]; break;
      case HB_OP_CONST: j_const(n, &nodes[k], o->c); break;
      case HB_OP_PARAM: if (o->a < 0 || o->a >= np) return -1; j_const(n, &nodes[k], prm[o->a]); break;
      case HB_OP_ADD: j_add(n, &nodes[k], &nodes[o->a], &nodes[o->b]); break;
      case HB_OP_SUB: j_sub(n, &nodes[k], &nodes[o->a], &nodes[o->b]); break;
      case HB_OP_MUL: j_mul(n, &nodes[k], &nodes[o->a], &nodes[o->b]); break;
      case HB_OP_DIV: j_div(n, &nodes[k], &nodes[o->a], &nodes[o->b]); break;
      case HB_OP_POW:
        if (ops[o->b].op == HB_OP_CONST) j_powc(n, &nodes[k], &nodes[o->a], ops[o->b].c);
        else j_pow(n, &nodes[k], &nodes[o->a], &nodes[o->b]);
        break;
      case HB_OP_POWI: j_powi(n, &nodes[k], &nodes[o->a], (int)o->c); break;
      case HB_OP_ATAN2: { /* atan2 y x: d = (x dy - y dx)/(x^2+y^2); via atan(y/x) derivatives, value fixed up */
        jet q; j_div(n, &q, &nodes[o->a], &nodes[o->b]); j_unary(n, &nodes[k], &q, HB_OP_ATAN);
        nodes[k].v = atan2(nodes[o->a].v, nodes[o->b].v); break; }
      default:
        if (o->op < 0 || o->op >= HB_OP__COUNT) return -1;
        j_unary(n, &nodes[k], &nodes[o->a], o->op);
    }
  }
  return 0;
}

/* logistic pos ht width x = ht / (1 + exp (-(beta * (x - pos)))), beta = log (0.9/(1-0.9)) / width
 * (app/Examples.hs:601-605) */
static void j_logistic(int n, jet* r, double pos, double ht, double width, const jet* x) {
  double beta = log(0.9 / (1 - 0.9)) / width;
  jet t; j_addc(n, &t, x, -pos); j_scale(n, &t, &t, beta); j_neg(n, &t, &t); j_exp(n, &t, &t);
  j_addc(n, &t, &t, 1.0); j_recip(n, &t, &t); j_scale(n, r, &t, ht);
}
static void j_one_minus(int n, jet* r, const jet* a) { jet t; j_neg(n, &t, a); j_addc(n, r, &t, 1.0); }

static const double BEZIER_DEFAULT[10] = {-1, -1, -2, 1, 0, 1, 1, -1, 2, 1}; /* app/Examples.hs:350 */

/* f: generalized -> Cartesian for the built-in fixtures */
static void builtin_f(const ho_system* S, const jet* q, jet* x) {
  int n = S->n; const double* P = S->prm; jet s, c, t;
  switch (S->builtin) {
    case HB_SYS_PENDULUM: /* app/Examples.hs:68: V2 (sin θ) (0.5 - cos θ) */
      j_sin(n, &x[0], &q[0]); j_cos(n, &c, &q[0]); j_neg(n, &c, &c); j_addc(n, &x[1], &c, 0.5); break;
    case HB_SYS_DOUBLE_PENDULUM: { /* app/Examples.hs:82-88 */
      jet s1, c1, s2, c2;
      j_sin(n, &s1, &q[0]); j_cos(n, &c1, &q[0]); j_sin(n, &s2, &q[1]); j_cos(n, &c2, &q[1]);
      x[0] = s1; j_one_minus(n, &x[1], &c1);
      j_scale(n, &t, &s2, 0.5); j_add(n, &x[2], &s1, &t);
      j_scale(n, &t, &c2, 0.5); j_sub(n, &x[3], &x[1], &t); break; }
    case HB_SYS_ROOM: x[0] = q[0]; x[1] = q[1]; break; /* id, app/Examples.hs:103 */
    case HB_SYS_TWO_BODY: { /* app/Examples.hs:129-136 */
      double m1 = P[0], m2 = P[1], mT = m1 + m2; jet r1, r2;
      j_scale(n, &r1, &q[0], -(m2 / mT)); j_scale(n, &r2, &q[0], m1 / mT);
      j_sin(n, &s, &q[1]); j_cos(n, &c, &q[1]);
      j_mul(n, &x[0], &r1, &c); j_mul(n, &x[1], &r1, &s); j_mul(n, &x[2], &r2, &c); j_mul(n, &x[3], &r2, &s); break; }
    case HB_SYS_SPRING: { /* app/Examples.hs:152: V3 r (r + (1+x) sin θ) ((1+x)(-cos θ)) */
      jet ox; j_addc(n, &ox, &q[1], 1.0); j_sin(n, &s, &q[2]); j_cos(n, &c, &q[2]); j_neg(n, &c, &c);
      x[0] = q[0]; j_mul(n, &t, &ox, &s); j_add(n, &x[1], &q[0], &t); j_mul(n, &x[2], &ox, &c); break; }
    case HB_SYS_BEZIER: { /* bezierCurve, app/Examples.hs:607-627, 5 control points => n' = 4 */
      static const int choose4[5] = {1, 4, 6, 4, 1};
      jet omt; j_one_minus(n, &omt, &q[0]);
      j_const(n, &x[0], 0.0); j_const(n, &x[1], 0.0);
      for (int i = 0; i < 5; i++) {
        jet a, b, coef; j_powi(n, &a, &omt, 4 - i); j_powi(n, &b, &q[0], i);
        j_scale(n, &coef, &a, (double)choose4[i]); j_mul(n, &coef, &coef, &b);
        j_scale(n, &t, &coef, P[2 * i]); j_add(n, &x[0], &x[0], &t);
        j_scale(n, &t, &coef, P[2 * i + 1]); j_add(n, &x[1], &x[1], &t);
      } break; }
    case HB_SYS_TRIPLE_PENDULUM: case HB_SYS_CHAIN12: { /* x_k = Σ l_i sin θ_i, y_k = 1 - Σ l_i cos θ_i */
      jet sx, sy; j_const(n, &sx, 0.0); j_const(n, &sy, 1.0);
      for (int k = 0; k < n; k++) {
        double l = (S->builtin == HB_SYS_TRIPLE_PENDULUM) ? P[3 + k] : 1.0;
        j_sin(n, &s, &q[k]); j_cos(n, &c, &q[k]);
        j_scale(n, &t, &s, l); j_add(n, &sx, &sx, &t);
        j_scale(n, &t, &c, l); j_sub(n, &sy, &sy, &t);
        x[2 * k] = sx; x[2 * k + 1] = sy;
      } break; }
    case HB_SYS_SPRING1D: /* synthetic: f x = (x cos α, x sin α) */
      j_scale(n, &x[0], &q[0], cos(P[1])); j_scale(n, &x[1], &q[0], sin(P[1])); break;
  }
}
/* potential; `in` = Cartesian jets if S->u_on_cart (mkSystem'), else generalized jets */
static void builtin_u(const ho_system* S, const jet* in, jet* u) {
  int n = S->n; const double* P = S->prm; jet t, a;
  switch (S->builtin) {
    case HB_SYS_PENDULUM: *u = in[1]; break; /* \(V2 _ y) -> y, app/Examples.hs:69 */
    case HB_SYS_DOUBLE_PENDULUM: /* 5 * (m1*y1 + m2*y2), app/Examples.hs:89 */
      j_scale(n, &t, &in[1], P[0]); j_scale(n, &a, &in[3], P[1]); j_add(n, &t, &t, &a); j_scale(n, u, &t, 5.0); break;
    case HB_SYS_ROOM: { /* app/Examples.hs:104-111, sum = foldl (+) 0 */
      jet acc; j_scale(n, &acc, &in[1], 2.0);
      j_logistic(n, &t, -1, 10, 0.1, &in[1]); j_one_minus(n, &t, &t); j_add(n, &acc, &acc, &t);
      j_logistic(n, &t, 1, 10, 0.1, &in[1]); j_add(n, &acc, &acc, &t);
      j_logistic(n, &t, -2, 10, 0.1, &in[0]); j_one_minus(n, &t, &t); j_add(n, &acc, &acc, &t);
      j_logistic(n, &t, 2, 10, 0.1, &in[0]); j_add(n, &acc, &acc, &t);
      *u = acc; break; }
    case HB_SYS_TWO_BODY: /* -(m1*m2 / r), app/Examples.hs:138 */
      j_recip(n, &t, &in[0]); j_scale(n, u, &t, -(P[0] * P[1])); break;
    case HB_SYS_SPRING: { /* app/Examples.hs:153-158 */
      jet acc, ox, c;
      j_powc(n, &t, &in[1], 2.0); j_scale(n, &t, &t, P[2]); j_scale(n, &acc, &t, 0.5);
      j_logistic(n, &t, -1.5, 25, 0.1, &in[0]); j_one_minus(n, &t, &t); j_add(n, &acc, &acc, &t);
      j_logistic(n, &t, 1.5, 25, 0.1, &in[0]); j_add(n, &acc, &acc, &t);
      j_addc(n, &ox, &in[1], 1.0); j_cos(n, &c, &in[2]); j_neg(n, &c, &c); j_mul(n, &t, &ox, &c);
      j_scale(n, &t, &t, P[0]); j_add(n, u, &acc, &t); break; }
    case HB_SYS_BEZIER: /* app/Examples.hs:176-179 */
      j_logistic(n, &t, 0, 5, 0.05, &in[0]); j_one_minus(n, &t, &t);
      j_logistic(n, &a, 1, 5, 0.05, &in[0]); j_add(n, u, &t, &a); break;
    case HB_SYS_TRIPLE_PENDULUM: case HB_SYS_CHAIN12: { /* 5 Σ m_k y_k */
      jet acc; j_const(n, &acc, 0.0);
      for (int k = 0; k < n; k++) {
        double mk = (S->builtin == HB_SYS_TRIPLE_PENDULUM) ? P[k] : 1.0;
        j_scale(n, &t, &in[2 * k + 1], mk); j_add(n, &acc, &acc, &t);
      }
      j_scale(n, u, &acc, 5.0); break; }
    case HB_SYS_SPRING1D: /* k x^2 / 2 */
      j_mul(n, &t, &in[0], &in[0]); j_scale(n, u, &t, 0.5 * P[0]); break;
  }
}

ho_system* ho_builtin(int id, const double* params, int np) {
  ho_system* S = (ho_system*)calloc(1, sizeof(ho_system));
  S->builtin = id;
  double* P = S->prm;
  #define SETP(k, dflt) P[k] = (params && np > (k)) ? params[k] : (dflt)
  switch (id) {
    case HB_SYS_PENDULUM: S->m = 2; S->n = 1; S->u_on_cart = 1; S->w[0] = S->w[1] = 1; break;
    case HB_SYS_DOUBLE_PENDULUM: S->m = 4; S->n = 2; S->u_on_cart = 1; S->np = 2; SETP(0, 1.0); SETP(1, 1.0);
      S->w[0] = S->w[1] = P[0]; S->w[2] = S->w[3] = P[1]; break;
    case HB_SYS_ROOM: S->m = 2; S->n = 2; S->w[0] = S->w[1] = 1; break;
    case HB_SYS_TWO_BODY: S->m = 4; S->n = 2; S->np = 2; SETP(0, 5.0); SETP(1, 0.5);
      S->w[0] = S->w[1] = P[0]; S->w[2] = S->w[3] = P[1]; break;
    case HB_SYS_SPRING: S->m = 3; S->n = 3; S->np = 3; SETP(0, 2.0); SETP(1, 1.0); SETP(2, 10.0);
      S->w[0] = P[0]; S->w[1] = S->w[2] = P[1]; break;
    case HB_SYS_BEZIER: S->m = 2; S->n = 1; S->np = 10; for (int k = 0; k < 10; k++) SETP(k, BEZIER_DEFAULT[k]);
      S->w[0] = S->w[1] = 1; break;
    case HB_SYS_TRIPLE_PENDULUM: S->m = 6; S->n = 3; S->u_on_cart = 1; S->np = 6;
      SETP(0, 1.0); SETP(1, 1.0); SETP(2, 1.0); SETP(3, 1.0); SETP(4, 0.5); SETP(5, 0.5);
      for (int k = 0; k < 3; k++) S->w[2 * k] = S->w[2 * k + 1] = P[k]; break;
    case HB_SYS_CHAIN12: S->m = 24; S->n = 12; S->u_on_cart = 1; for (int i = 0; i < 24; i++) S->w[i] = 1; break;
    case HB_SYS_SPRING1D: S->m = 2; S->n = 1; S->np = 2; SETP(0, 10.0); SETP(1, 0.3); S->w[0] = S->w[1] = 1; break;
    default: free(S); return NULL;
  }
  #undef SETP
  return S;
}

ho_system* ho_from_tape(int m, int n, const double* inertia, const hb_tape* f, const hb_tape* u,
                        int u_on_cart, const double* params, int np) {
  if (m < 1 || m > MMAX || n < 1 || n > NMAX || !f || !u || f->n_out != m || u->n_out != 1 ||
      f->n_in != n || u->n_in != (u_on_cart ? m : n) || np < 0 || np > HB_MAX_PARAMS) return NULL;
  ho_system* S = (ho_system*)calloc(1, sizeof(ho_system));
  S->builtin = -1; S->m = m; S->n = n; S->u_on_cart = u_on_cart; S->np = np;
  memcpy(S->w, inertia, sizeof(double) * m);
  if (np) memcpy(S->prm, params, sizeof(double) * np);
  S->nf = f->n_ops; S->nu = u->n_ops;
  S->f = (hb_op*)malloc(sizeof(hb_op) * (S->nf ? S->nf : 1)); memcpy(S->f, f->ops, sizeof(hb_op) * S->nf);
  S->u = (hb_op*)malloc(sizeof(hb_op) * (S->nu ? S->nu : 1)); memcpy(S->u, u->ops, sizeof(hb_op) * S->nu);
  for (int i = 0; i < m; i++) S->fout[i] = f->outs[i];
  S->uout = u->outs[0];
  return S;
}
void ho_free(ho_system* S) { if (S) { free(S->f); free(S->u); free(S); } }
void ho_dims(const ho_system* S, int* m, int* n) { *m = S->m; *n = S->n; }

/* x = f(q) as jets, and U as a jet over q (composition u . f when mkSystem') */
static int sys_eval(const ho_system* S, const double* q, jet* x, jet* U) {
  int n = S->n, m = S->m; jet qj[NMAX];
  for (int j = 0; j < n; j++) j_var(n, &qj[j], q[j], j);
  if (S->builtin >= 0) {
    builtin_f(S, qj, x);
    if (U) builtin_u(S, S->u_on_cart ? x : qj, U);
    return 0;
  }
  jet* nodes = (jet*)malloc(sizeof(jet) * (size_t)(S->nf > S->nu ? S->nf : S->nu));
  int rc = eval_tape(n, S->f, S->nf, qj, n, S->prm, S->np, nodes);
  if (!rc) for (int i = 0; i < m; i++) x[i] = nodes[S->fout[i]];
  if (!rc && U) {
    rc = eval_tape(n, S->u, S->nu, S->u_on_cart ? x : qj, S->u_on_cart ? m : n, S->prm, S->np, nodes);
    if (!rc) *U = nodes[S->uout];
  }
  free(nodes);
  return rc;
}

/* _sysCoords / underlyingPos (src/Numeric/Hamilton.hs:174-178, :220) */
int ho_underlying_pos(const ho_system* S, const double* q, double* x) {
  jet xj[MMAX]; int rc = sys_eval(S, q, xj, NULL);
  for (int i = 0; i < S->m; i++) x[i] = xj[i].v;
  return rc;
}
/* _sysPotential / pe (:182-186, :223) */
double ho_pe(const ho_system* S, const double* q) { jet xj[MMAX], U; sys_eval(S, q, xj, &U); return U.v; }
/* _sysJacobian (:221): J row-major m x n */
void ho_jacobian(const ho_system* S, const double* q, double* J) {
  jet xj[MMAX]; sys_eval(S, q, xj, NULL);
  for (int i = 0; i < S->m; i++) for (int j = 0; j < S->n; j++) J[i * S->n + j] = xj[i].g[j];
}
/* _sysHessian (:222, tr2 :227-233): H[j][i][k] = d2 f_i / dq_j dq_k, n slices of m x n */
void ho_hessian(const ho_system* S, const double* q, double* H) {
  jet xj[MMAX]; sys_eval(S, q, xj, NULL); int m = S->m, n = S->n;
  for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) for (int k = 0; k < n; k++) H[(j * m + i) * n + k] = xj[i].h[j][k];
}
/* _sysPotentialGrad (:224) */
void ho_potential_grad(const ho_system* S, const double* q, double* g) {
  jet xj[MMAX], U; sys_eval(S, q, xj, &U); for (int j = 0; j < S->n; j++) g[j] = U.g[j];
}

/* ----------------------------------------------------------- tiny dense linear algebra ----- */
static void matvec(int r, int c, const double* A, const double* x, double* y) { /* A r x c row-major */
  for (int i = 0; i < r; i++) { double s = 0; for (int j = 0; j < c; j++) s += A[i * c + j] * x[j]; y[i] = s; }
}
static void matTvec(int r, int c, const double* A, const double* x, double* y) { /* y = A^T x */
  for (int j = 0; j < c; j++) { double s = 0; for (int i = 0; i < r; i++) s += A[i * c + j] * x[i]; y[j] = s; }
}
/* inverse by Gauss-Jordan with partial pivoting (hmatrix `inv` = LAPACK LU solve against I) */
static int mat_inv(int n, const double* A, double* Ai) {
  double a[NMAX][2 * NMAX];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { a[i][j] = A[i * n + j]; a[i][n + j] = (i == j); }
  for (int c = 0; c < n; c++) {
    int piv = c; for (int r = c + 1; r < n; r++) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (a[piv][c] == 0.0 || !isfinite(a[piv][c])) return 1;
    if (piv != c) for (int j = 0; j < 2 * n; j++) { double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
    double d = 1.0 / a[c][c];
    for (int j = 0; j < 2 * n; j++) a[c][j] *= d;
    for (int r = 0; r < n; r++) if (r != c) { double f = a[r][c]; if (f != 0) for (int j = 0; j < 2 * n; j++) a[r][j] -= f * a[c][j]; }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ai[i * n + j] = a[i][n + j];
  return 0;
}
/* jmj = tr j <> diag w <> j  (:324, :380) */
static void mass_matrix(int m, int n, const double* J, const double* w, double* M) {
  for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) {
    double s = 0; for (int i = 0; i < m; i++) s += J[i * n + a] * w[i] * J[i * n + b]; M[a * n + b] = s;
  }
}

/* momenta (:262-269): tr j #> diag w #> j #> v */
void ho_momenta(const ho_system* S, const double* q, const double* v, double* p) {
  double J[MMAX * NMAX], t[MMAX]; ho_jacobian(S, q, J);
  matvec(S->m, S->n, J, v, t); for (int i = 0; i < S->m; i++) t[i] *= S->w[i]; matTvec(S->m, S->n, J, t, p);
}
/* velocities (:316-324): inv jmj #> p */
int ho_velocities(const ho_system* S, const double* q, const double* p, double* v) {
  double J[MMAX * NMAX], M[NMAX * NMAX], Mi[NMAX * NMAX]; ho_jacobian(S, q, J);
  mass_matrix(S->m, S->n, J, S->w, M); if (mat_inv(S->n, M, Mi)) return 1;
  matvec(S->n, S->n, Mi, p, v); return 0;
}
static double dot(int n, const double* a, const double* b) { double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; }
double ho_keC(const ho_system* S, const double* q, const double* v) { double p[NMAX]; ho_momenta(S, q, v, p); return dot(S->n, v, p) / 2; } /* :288-296 */
double ho_keP(const ho_system* S, const double* q, const double* p) { double v[NMAX]; if (ho_velocities(S, q, p, v)) return NAN; return dot(S->n, v, p) / 2; } /* :341-349 */
double ho_lagrangian(const ho_system* S, const double* q, const double* v) { return ho_keC(S, q, v) - ho_pe(S, q); }   /* :301-309 */
double ho_hamiltonian(const ho_system* S, const double* q, const double* p) { return ho_keP(S, q, p) + ho_pe(S, q); } /* :353-361 */

/* hamEqs (:370-387).  dTdq_j = -(p <.> ijmj #> trj #> mm #> djdq_j #> ijmj #> p), all #> right-associated. */
int ho_ham_eqs(const ho_system* S, const double* q, const double* p, double* dq, double* dp) {
  int m = S->m, n = S->n; jet xj[MMAX], U;
  if (sys_eval(S, q, xj, &U)) return 2;
  double J[MMAX * NMAX], M[NMAX * NMAX], Mi[NMAX * NMAX];
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) J[i * n + j] = xj[i].g[j];
  mass_matrix(m, n, J, S->w, M);
  if (mat_inv(n, M, Mi)) return 1;
  for (int j = 0; j < n; j++) {
    double Hj[MMAX * NMAX], a[NMAX], b[MMAX], c[NMAX], d[NMAX];
    for (int i = 0; i < m; i++) for (int k = 0; k < n; k++) Hj[i * n + k] = xj[i].h[j][k];
    matvec(n, n, Mi, p, a);            /* ijmj #> p            */
    matvec(m, n, Hj, a, b);            /* djdq #> ...          */
    for (int i = 0; i < m; i++) b[i] *= S->w[i]; /* mm #> ...  */
    matTvec(m, n, J, b, c);            /* trj #> ...           */
    matvec(n, n, Mi, c, d);            /* ijmj #> ...          */
    double dTdq = -dot(n, p, d);
    dp[j] = -(dTdq + U.g[j]);          /* -dHdq, dHdq = dTdq + gradU (:387, :375) */
  }
  matvec(n, n, Mi, p, dq);             /* dHdp (:386) */
  return 0;
}

/* ----------------------------------------------------------------------- integrators ------- */
typedef struct { long rhs_evals, steps, rejects; } ho_stats;
static int rhs(const ho_system* S, const double* y, double* dy, ho_stats* st) {
  if (st) st->rhs_evals++;
  return ho_ham_eqs(S, y, y + S->n, dy, dy + S->n);
}

/* classical RK4, `nsteps` steps of size dt on y = [q, p] (in place) */
int ho_rk4(const ho_system* S, double dt, int nsteps, double* y) {
  int d = 2 * S->n; double k1[2 * NMAX], k2[2 * NMAX], k3[2 * NMAX], k4[2 * NMAX], t[2 * NMAX];
  for (int s = 0; s < nsteps; s++) {
    if (rhs(S, y, k1, NULL)) return 1;
    for (int i = 0; i < d; i++) t[i] = y[i] + 0.5 * dt * k1[i];
    if (rhs(S, t, k2, NULL)) return 1;
    for (int i = 0; i < d; i++) t[i] = y[i] + 0.5 * dt * k2[i];
    if (rhs(S, t, k3, NULL)) return 1;
    for (int i = 0; i < d; i++) t[i] = y[i] + dt * k3[i];
    if (rhs(S, t, k4, NULL)) return 1;
    for (int i = 0; i < d; i++) y[i] += dt / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
  }
  return 0;
}

/* GSL rkf45 stepper (ode-initval2/rkf45.c): 5th-order solution advanced, |4th-5th| as yerr */
static int rkf45_apply(const ho_system* S, int d, double h, double* y, double* yerr,
                       const double* dydt_in, double* dydt_out, ho_stats* st) {
  static const double ah[] = {1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0};
  static const double b3[] = {3.0 / 32.0, 9.0 / 32.0};
  static const double b4[] = {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0};
  static const double b5[] = {8341.0 / 4104.0, -32832.0 / 4104.0, 29440.0 / 4104.0, -845.0 / 4104.0};
  static const double b6[] = {-6080.0 / 20520.0, 41040.0 / 20520.0, -28352.0 / 20520.0, 9295.0 / 20520.0, -5643.0 / 20520.0};
  static const double c1 = 902880.0 / 7618050.0, c3 = 3953664.0 / 7618050.0, c4 = 3855735.0 / 7618050.0,
                      c5 = -1371249.0 / 7618050.0, c6 = 277020.0 / 7618050.0;
  static const double ec[] = {0.0, 1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0, 2.0 / 55.0};
  double k1[2 * NMAX], k2[2 * NMAX], k3[2 * NMAX], k4[2 * NMAX], k5[2 * NMAX], k6[2 * NMAX], yt[2 * NMAX];
  (void)ah; /* autonomous system: stage times unused (`const f`, src/Numeric/Hamilton.hs:445) */
  memcpy(k1, dydt_in, sizeof(double) * d);
  for (int i = 0; i < d; i++) yt[i] = y[i] + ah[0] * h * k1[i];
  if (rhs(S, yt, k2, st)) return 1;
  for (int i = 0; i < d; i++) yt[i] = y[i] + h * (b3[0] * k1[i] + b3[1] * k2[i]);
  if (rhs(S, yt, k3, st)) return 1;
  for (int i = 0; i < d; i++) yt[i] = y[i] + h * (b4[0] * k1[i] + b4[1] * k2[i] + b4[2] * k3[i]);
  if (rhs(S, yt, k4, st)) return 1;
  for (int i = 0; i < d; i++) yt[i] = y[i] + h * (b5[0] * k1[i] + b5[1] * k2[i] + b5[2] * k3[i] + b5[3] * k4[i]);
  if (rhs(S, yt, k5, st)) return 1;
  for (int i = 0; i < d; i++) yt[i] = y[i] + h * (b6[0] * k1[i] + b6[1] * k2[i] + b6[2] * k3[i] + b6[3] * k4[i] + b6[4] * k5[i]);
  if (rhs(S, yt, k6, st)) return 1;
  for (int i = 0; i < d; i++) { const double di = c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i] + c6 * k6[i]; y[i] += h * di; }
  if (rhs(S, y, dydt_out, st)) return 1;
  for (int i = 0; i < d; i++) yerr[i] = h * (ec[1] * k1[i] + ec[3] * k3[i] + ec[4] * k4[i] + ec[5] * k5[i] + ec[6] * k6[i]);
  return 0;
}

/* GSL standard controller (cstd.c std_control_hadjust), a_y = a_dydt = 1 (hmatrix-gsl odeSolveV),
 * ord = 5 (rkf45_order).  returns -1 DEC, +1 INC, 0 NIL */
static int std_hadjust(int d, double eps_abs, double eps_rel, const double* y, const double* yerr, const double* yp, double* h) {
  const double a_y = 1.0, a_dydt = 1.0, S = 0.9, ord = 5.0, h_old = *h;
  double rmax = DBL_MIN;
  for (int i = 0; i < d; i++) {
    const double D0 = eps_rel * (a_y * fabs(y[i]) + a_dydt * fabs(h_old * yp[i])) + eps_abs;
    const double r = fabs(yerr[i]) / fabs(D0);
    rmax = r > rmax ? r : rmax;   /* GSL_MAX_DBL */
  }
  if (rmax > 1.1) { double r = S / pow(rmax, 1.0 / ord); if (r < 0.2) r = 0.2; *h = r * h_old; return -1; }
  else if (rmax < 0.5) { double r = S / pow(rmax, 1.0 / (ord + 1.0)); if (r > 5.0) r = 5.0; if (r < 1.0) r = 1.0; *h = r * h_old; return 1; }
  return 0;
}

typedef struct { long count; double dydt_in[2 * NMAX], dydt_out[2 * NMAX]; } evolve_state;

/* gsl_odeiv2_evolve_apply (evolve.c) for a stepper with can_use_dydt_in = 1 */
static int evolve_apply(const ho_system* S, evolve_state* e, int d, double eps, double* t, double t1, double* h, double* y, ho_stats* st) {
  const double t0 = *t; double h0 = *h; int final_step = 0; double dt = t1 - t0;
  double y0[2 * NMAX], yerr[2 * NMAX];
  memcpy(y0, y, sizeof(double) * d);
  if (e->count == 0) { if (rhs(S, y, e->dydt_in, st)) return 1; }
  else memcpy(e->dydt_in, e->dydt_out, sizeof(double) * d);
  for (;;) {
    if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) { h0 = dt; final_step = 1; } else final_step = 0;
    if (rkf45_apply(S, d, h0, y, yerr, e->dydt_in, e->dydt_out, st)) return 1;
    e->count++;
    if (final_step) *t = t1; else *t = t0 + h0;
    double h_old = h0;
    int adj = std_hadjust(d, eps, eps, y, yerr, e->dydt_out, &h0);
    if (adj == -1) {
      volatile double t_curr = *t, t_next = (*t) + h0;   /* GSL_COERCE_DBL */
      if (fabs(h0) < fabs(h_old) && t_next != t_curr) {
        memcpy(y, y0, sizeof(double) * d); if (st) st->rejects++; continue; /* undo, retry */
      } else { *h = h0; return 4; /* GSL_FAILURE */ }
    }
    break;
  }
  if (st) st->steps++;
  if (final_step == 0) *h = h0;
  return 0;
}

/* evolveHam (src/Numeric/Hamilton.hs:443-462) + hmatrix-gsl ode(): out is s x 2n, row 0 = initial state */
int ho_evolve_ham(const ho_system* S, const double* q0, const double* p0, const double* ts, int s, double* out, ho_stats* st) {
  int n = S->n, d = 2 * n; double y[2 * NMAX];
  const double eps = 1.49012e-08;             /* :448 */
  double h = (ts[1] - ts[0]) / 100;           /* hi, :447 */
  double t = ts[0];
  evolve_state e; e.count = 0;
  memcpy(y, q0, sizeof(double) * n); memcpy(y + n, p0, sizeof(double) * n);  /* fromPs :457-458 */
  memcpy(out, y, sizeof(double) * d);
  for (int i = 1; i < s; i++) {
    double ti = ts[i];
    while (t < ti) { int rc = evolve_apply(S, &e, d, eps, &t, ti, &h, y, st); if (rc) return rc; }
    memcpy(out + (size_t)i * d, y, sizeof(double) * d);
  }
  return 0;
}
/* stepHam (:400-402) */
int ho_step_ham(const ho_system* S, double r, const double* q, const double* p, double* qo, double* po, ho_stats* st) {
  double ts[2] = {0.0, r}, out[4 * NMAX]; int n = S->n;
  int rc = ho_evolve_ham(S, q, p, ts, 2, out, st);
  memcpy(qo, out + 2 * n, sizeof(double) * n); memcpy(po, out + 3 * n, sizeof(double) * n);
  return rc;
}

/* ------------------------------------------------------------------------ batch drivers ----
 * y is AOS N x 2n.  integ: 0 = RK4, 1 = RKF45_GSL (nsteps independent stepHam dt each).
 * pthreads over contiguous blocks of trajectories; returns the number of failed trajectories. */
#include <pthread.h>
typedef struct { const ho_system* S; int integ, nsteps; double dt; long lo, hi; double* y; const double* yin; long bad; } job_t;
static void* step_worker(void* arg) {
  job_t* J = (job_t*)arg; const ho_system* S = J->S; int d = 2 * S->n, n = S->n;
  for (long i = J->lo; i < J->hi; i++) {
    double* yi = J->y + i * d; int rc = 0;
    if (J->integ == 0) rc = ho_rk4(S, J->dt, J->nsteps, yi);
    else for (int s = 0; s < J->nsteps && !rc; s++) {
      double qo[NMAX], po[NMAX]; rc = ho_step_ham(S, J->dt, yi, yi + n, qo, po, NULL);
      memcpy(yi, qo, sizeof(double) * n); memcpy(yi + n, po, sizeof(double) * n);
    }
    if (rc) J->bad++;
  }
  return NULL;
}
static void* eqs_worker(void* arg) {
  job_t* J = (job_t*)arg; const ho_system* S = J->S; int d = 2 * S->n, n = S->n;
  for (long i = J->lo; i < J->hi; i++) ho_ham_eqs(S, J->yin + i * d, J->yin + i * d + n, J->y + i * d, J->y + i * d + n);
  return NULL;
}
static long run_jobs(void* (*fn)(void*), job_t proto, long N, int nthreads) {
  if (nthreads < 1) nthreads = 1; if (nthreads > 1024) nthreads = 1024; if ((long)nthreads > N) nthreads = N > 0 ? (int)N : 1;
  job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t)); pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) { jobs[t] = proto; jobs[t].lo = N * t / nthreads; jobs[t].hi = N * (t + 1) / nthreads; jobs[t].bad = 0; }
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, fn, &jobs[t]);
  fn(&jobs[0]);
  long bad = jobs[0].bad;
  for (int t = 1; t < nthreads; t++) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
  free(jobs); free(th); return bad;
}
long ho_batch_step(const ho_system* S, int integ, double dt, int nsteps, long N, double* y, int nthreads) {
  job_t p; memset(&p, 0, sizeof p); p.S = S; p.integ = integ; p.nsteps = nsteps; p.dt = dt; p.y = y;
  return run_jobs(step_worker, p, N, nthreads);
}
void ho_batch_ham_eqs(const ho_system* S, long N, const double* y, double* dy, int nthreads) {
  job_t p; memset(&p, 0, sizeof p); p.S = S; p.yin = y; p.y = dy;
  run_jobs(eqs_worker, p, N, nthreads);
}
/* splitmix64 initial Phases, AOS */
void ho_init_random(const ho_system* S, uint64_t seed, long first, long N, const double* lo, const double* hi, double* y) {
  int d = 2 * S->n;
  for (long i = 0; i < N; i++) for (int c = 0; c < d; c++) {
    uint64_t z = seed + (uint64_t)d * (uint64_t)(first + i) + (uint64_t)c;
    z += 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    y[i * d + c] = lo[c] + (hi[c] - lo[c]) * u;
  }
}
#include <unistd.h>
int ho_max_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
