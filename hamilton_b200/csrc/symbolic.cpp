// symbolic.cpp — expression DAG with folding + symbolic second-order forward-mode jets.
#include "symbolic.hpp"

#include <cstring>

namespace hb {

static uint64_t bits(double c) { uint64_t u; std::memcpy(&u, &c, 8); return u; }

int Graph::intern(const Node& n) {
  Key k{(uint8_t)n.op, n.a, n.b, bits(n.c)};
  auto it = cse_.find(k);
  if (it != cse_.end()) return it->second;
  nodes.push_back(n);
  int id = (int)nodes.size() - 1;
  cse_.emplace(k, id);
  return id;
}

int Graph::constant(double c) {
  if (c == 0.0) c = 0.0;  // canonical +0
  return intern({Op::Const, -1, -1, c});
}

// a = coef * base with a literal coefficient (1 when a is not such a product)
static void split_coef(const Graph& G, int a, double* coef, int* base) {
  double c;
  if (G.nodes[a].op == Op::Mul && G.is_const(G.nodes[a].a, &c)) { *coef = c; *base = G.nodes[a].b; }
  else { *coef = 1.0; *base = a; }
}

int Graph::add(int a, int b) {
  double x, y;
  if (is_const(a, &x) && is_const(b, &y)) return constant(x + y);
  if (is_zero(a)) return b;
  if (is_zero(b)) return a;
  if (nodes[b].op == Op::Neg) return sub(a, nodes[b].a);
  if (nodes[a].op == Op::Neg) return sub(b, nodes[a].a);
  if (a == b) return mul(constant(2.0), a);   // x + x -> 2 x (merges with neighbouring constant factors)
  {   // like terms: c1 x + c2 x -> (c1 + c2) x  (e.g. the gradient of m g (y1 + y2 + y3) of a pendulum chain: s + 2 s -> 3 s)
    double ca, cb; int ba, bb;
    split_coef(*this, a, &ca, &ba); split_coef(*this, b, &cb, &bb);
    if (ba == bb && !is_const(ba)) return mul(constant(ca + cb), ba);
  }
  if (a > b) std::swap(a, b);
  return intern({Op::Add, a, b, 0.0});
}

int Graph::sub(int a, int b) {
  double x, y;
  if (is_const(a, &x) && is_const(b, &y)) return constant(x - y);
  if (is_zero(b)) return a;
  if (is_zero(a)) return neg(b);
  if (a == b) return constant(0.0);
  if (nodes[b].op == Op::Neg) return add(a, nodes[b].a);
  if (nodes[a].op == Op::Neg) return neg(add(nodes[a].a, b));
  {   // like terms: c1 x - c2 x -> (c1 - c2) x
    double ca, cb; int ba, bb;
    split_coef(*this, a, &ca, &ba); split_coef(*this, b, &cb, &bb);
    if (ba == bb && !is_const(ba)) return mul(constant(ca - cb), ba);
  }
  return intern({Op::Sub, a, b, 0.0});
}

int Graph::mul(int a, int b) {
  double x, y;
  if (is_const(a, &x) && is_const(b, &y)) return constant(x * y);
  if (is_zero(a) || is_zero(b)) return constant(0.0);
  if (is_one(a)) return b;
  if (is_one(b)) return a;
  if (is_const(a, &x) && x == -1.0) return neg(b);
  if (is_const(b, &y) && y == -1.0) return neg(a);
  if (nodes[a].op == Op::Neg) return neg(mul(nodes[a].a, b));
  if (nodes[b].op == Op::Neg) return neg(mul(a, nodes[b].a));
  if (is_const(b)) std::swap(a, b);  // constants first
  if (is_const(a, &x) && nodes[b].op == Op::Mul && is_const(nodes[b].a, &y)) return mul(constant(x * y), nodes[b].b);
  if (!is_const(a) && a > b) std::swap(a, b);
  return intern({Op::Mul, a, b, 0.0});
}

int Graph::neg(int a) {
  double x;
  if (is_const(a, &x)) return constant(-x);
  if (nodes[a].op == Op::Neg) return nodes[a].a;
  if (nodes[a].op == Op::Sub) return sub(nodes[a].b, nodes[a].a);
  if (nodes[a].op == Op::Mul && is_const(nodes[a].a, &x)) return mul(constant(-x), nodes[a].b);
  return intern({Op::Neg, a, -1, 0.0});
}

int Graph::recip(int a) {
  double x;
  if (is_const(a, &x)) return constant(1.0 / x);
  if (nodes[a].op == Op::Recip) return nodes[a].a;
  if (nodes[a].op == Op::Neg) return neg(recip(nodes[a].a));
  return intern({Op::Recip, a, -1, 0.0});
}

static double fold_unary(Op op, double x) {
  switch (op) {
    case Op::Abs: return std::fabs(x);
    case Op::Signum: return (double)((x > 0) - (x < 0));
    case Op::Sqrt: return std::sqrt(x);
    case Op::Exp: return std::exp(x);
    case Op::Log: return std::log(x);
    case Op::Sin: return std::sin(x);
    case Op::Cos: return std::cos(x);
    case Op::Tan: return std::tan(x);
    case Op::Asin: return std::asin(x);
    case Op::Acos: return std::acos(x);
    case Op::Atan: return std::atan(x);
    case Op::Sinh: return std::sinh(x);
    case Op::Cosh: return std::cosh(x);
    case Op::Tanh: return std::tanh(x);
    case Op::Asinh: return std::asinh(x);
    case Op::Acosh: return std::acosh(x);
    case Op::Atanh: return std::atanh(x);
    default: return NAN;
  }
}

int Graph::unary(Op op, int a) {
  double x;
  if (op == Op::Neg) return neg(a);
  if (op == Op::Recip) return recip(a);
  if (is_const(a, &x)) return constant(fold_unary(op, x));
  if (nodes[a].op == Op::Neg) {  // parity: keeps sin(-x)/cos(-x) sharing the sincos of x
    switch (op) {
      case Op::Sin: case Op::Tan: case Op::Asin: case Op::Atan: case Op::Sinh: case Op::Tanh: case Op::Asinh: case Op::Atanh:
      case Op::Signum:
        return neg(unary(op, nodes[a].a));
      case Op::Cos: case Op::Cosh: case Op::Abs:
        return unary(op, nodes[a].a);
      default: break;
    }
  }
  return intern({op, a, -1, 0.0});
}

int Graph::pow(int a, int b) {
  double x, y;
  if (is_const(b, &y)) {
    if (is_const(a, &x)) return constant(std::pow(x, y));
    if (y == 0.0) return constant(1.0);
    if (y == 1.0) return a;
    if (y == 2.0) return mul(a, a);
    if (y == -1.0) return recip(a);
    if (y == 0.5) return unary(Op::Sqrt, a);
  }
  return intern({Op::Pow, a, b, 0.0});
}

int Graph::atan2(int a, int b) {
  double x, y;
  if (is_const(a, &x) && is_const(b, &y)) return constant(std::atan2(x, y));
  return intern({Op::Atan2, a, b, 0.0});
}

// ------------------------------------------------------------------------------- jets -------
void JetAlgebra::acc(std::map<int, int>& m, int key, int node) {
  if (G.is_zero(node)) return;
  auto it = m.find(key);
  if (it == m.end()) { m.emplace(key, node); return; }
  int s = G.add(it->second, node);
  if (G.is_zero(s)) m.erase(it); else it->second = s;
}
void JetAlgebra::acc(std::map<std::pair<int, int>, int>& m, std::pair<int, int> key, int node) {
  if (G.is_zero(node)) return;
  auto it = m.find(key);
  if (it == m.end()) { m.emplace(key, node); return; }
  int s = G.add(it->second, node);
  if (G.is_zero(s)) m.erase(it); else it->second = s;
}

SJet JetAlgebra::add(const SJet& a, const SJet& b) {
  SJet r = a;
  r.v = G.add(a.v, b.v);
  for (auto& kv : b.g) acc(r.g, kv.first, kv.second);
  if (order_ >= 2) for (auto& kv : b.h) acc(r.h, kv.first, kv.second);
  return r;
}
SJet JetAlgebra::neg(const SJet& a) {
  SJet r;
  r.v = G.neg(a.v);
  for (auto& kv : a.g) r.g[kv.first] = G.neg(kv.second);
  if (order_ >= 2) for (auto& kv : a.h) r.h[kv.first] = G.neg(kv.second);
  return r;
}
SJet JetAlgebra::sub(const SJet& a, const SJet& b) {
  SJet r = a;
  r.v = G.sub(a.v, b.v);
  for (auto& kv : b.g) acc(r.g, kv.first, G.neg(kv.second));
  if (order_ >= 2) for (auto& kv : b.h) acc(r.h, kv.first, G.neg(kv.second));
  return r;
}
SJet JetAlgebra::mul(const SJet& a, const SJet& b) {
  SJet r;
  r.v = G.mul(a.v, b.v);
  for (auto& kv : b.g) acc(r.g, kv.first, G.mul(a.v, kv.second));
  for (auto& kv : a.g) acc(r.g, kv.first, G.mul(b.v, kv.second));
  if (order_ >= 2) {
    for (auto& kv : b.h) acc(r.h, kv.first, G.mul(a.v, kv.second));
    for (auto& kv : a.h) acc(r.h, kv.first, G.mul(b.v, kv.second));
    for (auto& ga : a.g)
      for (auto& gb : b.g) {
        int j = ga.first, k = gb.first;
        int t = G.mul(ga.second, gb.second);
        if (j == k) acc(r.h, {j, j}, G.scale(2.0, t));           // a_j b_j + a_j b_j
        else acc(r.h, {j < k ? j : k, j < k ? k : j}, t);        // a_j b_k contributes to (j,k); (k,j) arrives from the other pair
      }
  }
  return r;
}
SJet JetAlgebra::chain(const SJet& a, int f0, int f1, int f2) {
  SJet r;
  r.v = f0;
  for (auto& kv : a.g) acc(r.g, kv.first, G.mul(f1, kv.second));
  if (order_ >= 2) {
    for (auto& kv : a.h) acc(r.h, kv.first, G.mul(f1, kv.second));
    if (!G.is_zero(f2))
      for (auto i1 = a.g.begin(); i1 != a.g.end(); ++i1)
        for (auto i2 = i1; i2 != a.g.end(); ++i2)
          acc(r.h, {i1->first, i2->first}, G.mul(f2, G.mul(i1->second, i2->second)));
  }
  return r;
}
SJet JetAlgebra::recip(const SJet& a) {
  int r = G.recip(a.v), r2 = G.mul(r, r);
  return chain(a, r, G.neg(r2), G.scale(2.0, G.mul(r2, r)));
}
SJet JetAlgebra::powi(const SJet& a, int k) {
  // Num (^): repeated multiplication by squaring, as GHC's (^) does (bezierCurve, app/Examples.hs:618)
  if (k == 0) return constant(1.0);
  bool negk = k < 0;
  if (negk) k = -k;
  SJet base = a, accj;
  bool have = false;
  while (k) {
    if (k & 1) { accj = have ? mul(accj, base) : base; have = true; }
    k >>= 1;
    if (k) base = mul(base, base);
  }
  return negk ? recip(accj) : accj;
}
SJet JetAlgebra::pow(const SJet& a, const SJet& b) {
  double c;
  if (G.is_const(b.v, &c) && b.g.empty()) {
    // literal exponent: only the base is differentiated (ad's (**) with a known-constant exponent),
    // so negative bases with integral exponents stay finite (spring: x ** 2, app/Examples.hs:154)
    int f0 = G.pow(a.v, b.v);
    int f1 = G.scale(c, G.pow(a.v, G.constant(c - 1.0)));
    int f2 = G.scale(c * (c - 1.0), G.pow(a.v, G.constant(c - 2.0)));
    return chain(a, f0, f1, f2);
  }
  SJet l = unary(HB_OP_LOG, a);
  SJet e = unary(HB_OP_EXP, mul(l, b));
  e.v = G.pow(a.v, b.v);
  return e;
}
SJet JetAlgebra::atan2(const SJet& a, const SJet& b) {
  SJet r = unary(HB_OP_ATAN, div(a, b));
  r.v = G.atan2(a.v, b.v);
  return r;
}

SJet JetAlgebra::unary(int opc, const SJet& a) {
  const int x = a.v;
  auto c = [&](double v) { return G.constant(v); };
  switch (opc) {
    case HB_OP_NEG: return neg(a);
    case HB_OP_RECIP: return recip(a);
    case HB_OP_ABS: return chain(a, G.unary(Op::Abs, x), G.unary(Op::Signum, x), c(0.0));
    case HB_OP_SIGNUM: return leaf(G.unary(Op::Signum, x));
    case HB_OP_SQRT: {
      int s = G.unary(Op::Sqrt, x), r = G.recip(s);
      return chain(a, s, G.scale(0.5, r), G.scale(-0.25, G.mul(r, G.mul(r, r))));
    }
    case HB_OP_EXP: { int e = G.unary(Op::Exp, x); return chain(a, e, e, e); }
    case HB_OP_LOG: { int r = G.recip(x); return chain(a, G.unary(Op::Log, x), r, G.neg(G.mul(r, r))); }
    case HB_OP_SIN: { int s = G.unary(Op::Sin, x), co = G.unary(Op::Cos, x); return chain(a, s, co, G.neg(s)); }
    case HB_OP_COS: { int s = G.unary(Op::Sin, x), co = G.unary(Op::Cos, x); return chain(a, co, G.neg(s), G.neg(co)); }
    case HB_OP_TAN: {
      int t = G.unary(Op::Tan, x), u = G.add(c(1.0), G.mul(t, t));
      return chain(a, t, u, G.scale(2.0, G.mul(t, u)));
    }
    case HB_OP_ASIN: case HB_OP_ACOS: {
      int u = G.sub(c(1.0), G.mul(x, x)), r = G.recip(G.unary(Op::Sqrt, u)), r3 = G.mul(r, G.mul(r, r));
      if (opc == HB_OP_ASIN) return chain(a, G.unary(Op::Asin, x), r, G.mul(x, r3));
      return chain(a, G.unary(Op::Acos, x), G.neg(r), G.neg(G.mul(x, r3)));
    }
    case HB_OP_ATAN: {
      int r = G.recip(G.add(c(1.0), G.mul(x, x)));
      return chain(a, G.unary(Op::Atan, x), r, G.scale(-2.0, G.mul(x, G.mul(r, r))));
    }
    case HB_OP_SINH: { int s = G.unary(Op::Sinh, x), co = G.unary(Op::Cosh, x); return chain(a, s, co, s); }
    case HB_OP_COSH: { int s = G.unary(Op::Sinh, x), co = G.unary(Op::Cosh, x); return chain(a, co, s, co); }
    case HB_OP_TANH: {
      int t = G.unary(Op::Tanh, x), u = G.sub(c(1.0), G.mul(t, t));
      return chain(a, t, u, G.scale(-2.0, G.mul(t, u)));
    }
    case HB_OP_ASINH: case HB_OP_ACOSH: {
      int u = (opc == HB_OP_ASINH) ? G.add(G.mul(x, x), c(1.0)) : G.sub(G.mul(x, x), c(1.0));
      int r = G.recip(G.unary(Op::Sqrt, u)), r3 = G.mul(r, G.mul(r, r));
      return chain(a, G.unary(opc == HB_OP_ASINH ? Op::Asinh : Op::Acosh, x), r, G.neg(G.mul(x, r3)));
    }
    case HB_OP_ATANH: {
      int r = G.recip(G.sub(c(1.0), G.mul(x, x)));
      return chain(a, G.unary(Op::Atanh, x), r, G.scale(2.0, G.mul(x, G.mul(r, r))));
    }
    default: return leaf(G.constant(NAN));
  }
}

bool replay_tape(JetAlgebra& A, const hb_op* ops, int n_ops, const std::vector<SJet>& inputs, int n_params,
                 std::vector<SJet>& nodes, std::string& err, const double* bake) {
  nodes.assign((size_t)n_ops, SJet());
  auto bad = [&](int k, const char* what) { err = "tape node " + std::to_string(k) + ": " + what; return false; };
  for (int k = 0; k < n_ops; k++) {
    const hb_op& o = ops[k];
    const bool binary = (o.op >= HB_OP_ADD && o.op <= HB_OP_DIV) || o.op == HB_OP_POW || o.op == HB_OP_ATAN2;
    if (o.op < 0 || o.op >= HB_OP__COUNT) return bad(k, "unknown opcode");
    if (o.op >= HB_OP_ADD && (o.a < 0 || o.a >= k)) return bad(k, "operand a must refer to an earlier node");
    if (binary && (o.b < 0 || o.b >= k)) return bad(k, "operand b must refer to an earlier node");
    switch (o.op) {
      case HB_OP_INPUT:
        if (o.a < 0 || o.a >= (int)inputs.size()) return bad(k, "input index out of range");
        nodes[k] = inputs[o.a];
        break;
      case HB_OP_CONST: nodes[k] = A.constant(o.c); break;
      case HB_OP_PARAM:
        if (o.a < 0 || o.a >= n_params) return bad(k, "parameter index out of range");
        nodes[k] = bake ? A.constant(bake[o.a]) : A.leaf(A.G.param(o.a));
        break;
      case HB_OP_ADD: nodes[k] = A.add(nodes[o.a], nodes[o.b]); break;
      case HB_OP_SUB: nodes[k] = A.sub(nodes[o.a], nodes[o.b]); break;
      case HB_OP_MUL: nodes[k] = A.mul(nodes[o.a], nodes[o.b]); break;
      case HB_OP_DIV: nodes[k] = A.div(nodes[o.a], nodes[o.b]); break;
      case HB_OP_POW: nodes[k] = A.pow(nodes[o.a], nodes[o.b]); break;
      case HB_OP_POWI:
        if (o.c != std::floor(o.c) || std::fabs(o.c) > 1024) return bad(k, "POWI exponent must be a small integer");
        nodes[k] = A.powi(nodes[o.a], (int)o.c);
        break;
      case HB_OP_ATAN2: nodes[k] = A.atan2(nodes[o.a], nodes[o.b]); break;
      default: nodes[k] = A.unary(o.op, nodes[o.a]);
    }
  }
  return true;
}

}  // namespace hb
