// builtins.cpp — the reference's example systems (app/Examples.hs:61-183) and the two benchmark
// extensions as tapes, written against the tracing number type exactly the way
// the Haskell originals are written against `RealFloat a`.  Physical constants that the Haskell
// closes over (`realToFrac m1` ...) are runtime parameters here so that one ahead-of-time compiled
// kernel serves every parameter choice.
#include "builtins.hpp"

#include <cmath>

#include "../../include/hamilton_b200_trace.hpp"

namespace hb {
namespace {

// logistic pos ht width x = ht / (1 + exp (-(beta * (x - pos)))),  app/Examples.hs:601-605
Ex logistic(double pos, double ht, double width, const Ex& x) {
  const double beta = std::log(0.9 / (1 - 0.9)) / width;
  return ht / (1.0 + exp(-(beta * (x - pos))));
}

struct Builder {
  SystemSpec s;
  Tape tf, tu;
  std::vector<Ex> q, uin;
  void begin(int m, int n, int np, bool u_cart) {
    s.m = m; s.n = n; s.n_params = np; s.u_on_cartesian = u_cart;
    for (int j = 0; j < n; j++) q.push_back(Ex::input(tf, j));
    for (int j = 0; j < (u_cart ? m : n); j++) uin.push_back(Ex::input(tu, j));
  }
  void w_lit(double v) { InertiaTerm t; t.value = v; s.inertia.push_back(t); }
  void w_prm(int k) { InertiaTerm t; t.is_param = true; t.param = k; s.inertia.push_back(t); }
  Ex pf(int k) { return Ex::param(tf, k); }
  Ex pu(int k) { return Ex::param(tu, k); }
  void out_f(const std::vector<Ex>& x) { for (auto& e : x) s.f_outs.push_back(e.id); }
  SystemSpec finish(const Ex& u) { s.u_out = u.id; s.f_ops = tf.ops; s.u_ops = tu.ops; return s; }
};

}  // namespace

static const double BEZIER_DEFAULT[10] = {-1, -1, -2, 1, 0, 1, 1, -1, 2, 1};   // app/Examples.hs:350

const char* builtin_name(int id) {
  static const char* names[HB_SYS__COUNT] = {"pendulum", "double_pendulum", "room", "two_body", "spring",
                                             "bezier", "triple_pendulum", "chain12", "spring1d"};
  return (id >= 0 && id < HB_SYS__COUNT) ? names[id] : nullptr;
}

// user parameters (hb_system_builtin) -> tape parameters (HB_OP_PARAM values)
bool builtin_params(int id, const double* user, int n_user, std::vector<double>& out) {
  auto U = [&](int k, double dflt) { return (user && k < n_user) ? user[k] : dflt; };
  out.clear();
  switch (id) {
    case HB_SYS_PENDULUM: case HB_SYS_ROOM: case HB_SYS_CHAIN12: return true;
    case HB_SYS_DOUBLE_PENDULUM: out = {U(0, 1.0), U(1, 1.0)}; return true;
    case HB_SYS_TWO_BODY: {   // app/Examples.hs:121-131: mT = m1 + m2, r1 = r * (-(m2/mT)), r2 = r * (m1/mT)
      double m1 = U(0, 5.0), m2 = U(1, 0.5), mT = m1 + m2;
      out = {m1, m2, -(m2 / mT), m1 / mT, m1 * m2};
      return true;
    }
    case HB_SYS_SPRING: out = {U(0, 2.0), U(1, 1.0), U(2, 10.0)}; return true;
    case HB_SYS_BEZIER: for (int k = 0; k < 10; k++) out.push_back(U(k, BEZIER_DEFAULT[k])); return true;
    case HB_SYS_TRIPLE_PENDULUM: out = {U(0, 1.0), U(1, 1.0), U(2, 1.0), U(3, 1.0), U(4, 0.5), U(5, 0.5)}; return true;
    case HB_SYS_SPRING1D: { double k = U(0, 10.0), al = U(1, 0.3); out = {k, std::cos(al), std::sin(al)}; return true; }
  }
  return false;
}

bool builtin_spec(int id, SystemSpec& spec) {
  Builder b;
  switch (id) {
    case HB_SYS_PENDULUM: {   // app/Examples.hs:64-69
      b.begin(2, 1, 0, true); b.w_lit(1); b.w_lit(1);
      b.out_f({sin(b.q[0]), 0.5 - cos(b.q[0])});
      spec = b.finish(b.uin[1] + 0.0);
      return true;
    }
    case HB_SYS_DOUBLE_PENDULUM: {   // app/Examples.hs:78-89
      b.begin(4, 2, 2, true); b.w_prm(0); b.w_prm(0); b.w_prm(1); b.w_prm(1);
      const Ex &t1 = b.q[0], &t2 = b.q[1];
      b.out_f({sin(t1), 1.0 - cos(t1), sin(t1) + sin(t2) / 2.0, 1.0 - cos(t1) - cos(t2) / 2.0});
      spec = b.finish(5.0 * (b.pu(0) * b.uin[1] + b.pu(1) * b.uin[3]));
      return true;
    }
    case HB_SYS_ROOM: {   // app/Examples.hs:99-112
      b.begin(2, 2, 0, false); b.w_lit(1); b.w_lit(1);
      b.out_f({b.q[0] + 0.0, b.q[1] + 0.0});
      const Ex &x = b.uin[0], &y = b.uin[1];
      Ex u = 2.0 * y;
      u = u + (1.0 - logistic(-1, 10, 0.1, y));
      u = u + logistic(1, 10, 0.1, y);
      u = u + (1.0 - logistic(-2, 10, 0.1, x));
      u = u + logistic(2, 10, 0.1, x);
      spec = b.finish(u);
      return true;
    }
    case HB_SYS_TWO_BODY: {   // app/Examples.hs:123-138; params: m1, m2, -(m2/mT), m1/mT, m1*m2
      b.begin(4, 2, 5, false); b.w_prm(0); b.w_prm(0); b.w_prm(1); b.w_prm(1);
      const Ex &r = b.q[0], &th = b.q[1];
      Ex r1 = r * b.pf(2), r2 = r * b.pf(3);
      b.out_f({r1 * cos(th), r1 * sin(th), r2 * cos(th), r2 * sin(th)});
      spec = b.finish(-(b.pu(4) / b.uin[0]));
      return true;
    }
    case HB_SYS_SPRING: {   // app/Examples.hs:148-158; params: mB, mW, k
      b.begin(3, 3, 3, false); b.w_prm(0); b.w_prm(1); b.w_prm(1);
      {
        const Ex &r = b.q[0], &x = b.q[1], &th = b.q[2];
        b.out_f({r + 0.0, r + (1.0 + x) * sin(th), (1.0 + x) * (-cos(th))});
      }
      const Ex &r = b.uin[0], &x = b.uin[1], &th = b.uin[2];
      Ex u = b.pu(2) * pow(x, 2.0) / 2.0;
      u = u + (1.0 - logistic(-1.5, 25, 0.1, r));
      u = u + logistic(1.5, 25, 0.1, r);
      u = u + b.pu(0) * ((1.0 + x) * (-cos(th)));
      spec = b.finish(u);
      return true;
    }
    case HB_SYS_BEZIER: {   // app/Examples.hs:171-179, bezierCurve :607-627 with 5 control points (n' = 4)
      b.begin(2, 1, 10, false); b.w_lit(1); b.w_lit(1);
      static const int choose4[5] = {1, 4, 6, 4, 1};
      const Ex& t = b.q[0];
      Ex bx = Ex::constant(b.tf, 0.0), by = Ex::constant(b.tf, 0.0);
      for (int i = 0; i < 5; i++) {
        Ex coef = (double)choose4[i] * powi(1.0 - t, 4 - i) * powi(t, i);
        bx = bx + b.pf(2 * i) * coef;
        by = by + b.pf(2 * i + 1) * coef;
      }
      b.out_f({bx, by});
      const Ex& tu = b.uin[0];
      spec = b.finish((1.0 - logistic(0, 5, 0.05, tu)) + logistic(1, 5, 0.05, tu));
      return true;
    }
    case HB_SYS_TRIPLE_PENDULUM: case HB_SYS_CHAIN12: {
      // x_k = sum_{i<=k} l_i sin th_i, y_k = 1 - sum_{i<=k} l_i cos th_i, U = 5 sum m_k y_k  
      const bool triple = id == HB_SYS_TRIPLE_PENDULUM;
      const int n = triple ? 3 : 12;
      b.begin(2 * n, n, triple ? 6 : 0, true);
      for (int k = 0; k < n; k++) { if (triple) { b.w_prm(k); b.w_prm(k); } else { b.w_lit(1); b.w_lit(1); } }
      std::vector<Ex> xs;
      Ex sx = Ex::constant(b.tf, 0.0), sy = Ex::constant(b.tf, 1.0);
      for (int k = 0; k < n; k++) {
        if (triple) { sx = sx + b.pf(3 + k) * sin(b.q[k]); sy = sy - b.pf(3 + k) * cos(b.q[k]); }
        else { sx = sx + sin(b.q[k]); sy = sy - cos(b.q[k]); }
        xs.push_back(sx); xs.push_back(sy);
      }
      b.out_f(xs);
      Ex u = Ex::constant(b.tu, 0.0);
      for (int k = 0; k < n; k++) u = u + (triple ? b.pu(k) * b.uin[2 * k + 1] : b.uin[2 * k + 1]);
      spec = b.finish(5.0 * u);
      return true;
    }
    case HB_SYS_SPRING1D: {   // synthetic: f x = (x cos a, x sin a), U = k x^2 / 2; params: k, cos a, sin a
      b.begin(2, 1, 3, false); b.w_lit(1); b.w_lit(1);
      b.out_f({b.q[0] * b.pf(1), b.q[0] * b.pf(2)});
      spec = b.finish(b.pu(0) * (b.uin[0] * b.uin[0]) / 2.0);
      return true;
    }
  }
  return false;
}

}  // namespace hb
