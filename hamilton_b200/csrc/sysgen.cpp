// sysgen.cpp — prints the derivative DAG of a System as a `Sys` struct for the CUDA engine.
#include "sysgen.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <sstream>

#include "polyform.hpp"
#include "symbolic.hpp"

namespace hb {
namespace {

std::string lit(double c) {
  if (std::isnan(c)) return "__longlong_as_double(0x7ff8000000000000LL)";
  if (std::isinf(c)) return c > 0 ? "__longlong_as_double(0x7ff0000000000000LL)" : "__longlong_as_double(0xfff0000000000000LL)";
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.17g", c);
  std::string s(buf);
  if (s.find_first_of(".eE") == std::string::npos) s += ".0";
  if (c < 0 || (c == 0 && std::signbit(c))) s = "(" + s + ")";
  return s;
}

struct Emitter {
  const Graph& G;
  int n_q = 1 << 30;                  // Input indices >= n_q are the velocities v[idx - n_q] (hpost)
  std::map<int, std::string> ext;     // nodes computed elsewhere: printed by name, never descended into
  explicit Emitter(const Graph& g) : G(g) {}

  std::string name(int id) const {
    auto it = ext.find(id);
    if (it != ext.end()) return it->second;
    const Node& n = G.nodes[id];
    switch (n.op) {
      case Op::Const: return lit(n.c);
      case Op::Input: return n.a >= n_q ? "v[" + std::to_string(n.a - n_q) + "]" : "q[" + std::to_string(n.a) + "]";
      case Op::Param: return "prm[" + std::to_string(n.a) + "]";
      default: return "t" + std::to_string(id);
    }
  }

  // Straight-line code computing `outs` (lvalue, node).
  std::string body(const std::vector<std::pair<std::string, int>>& outs) const {
    std::vector<char> live(G.nodes.size(), 0);
    std::vector<int> stack;
    for (auto& o : outs) stack.push_back(o.second);
    while (!stack.empty()) {
      int id = stack.back(); stack.pop_back();
      if (live[id] || ext.count(id)) continue;
      live[id] = 1;
      const Node& n = G.nodes[id];
      if (n.op == Op::Const || n.op == Op::Input || n.op == Op::Param) continue;
      if (n.a >= 0) stack.push_back(n.a);
      if (n.b >= 0) stack.push_back(n.b);
    }
    // sin/cos of the same argument -> one sincos
    std::map<int, std::pair<int, int>> sc;   // arg -> (sin id, cos id)
    for (int id = 0; id < (int)G.nodes.size(); id++) {
      if (!live[id]) continue;
      const Node& n = G.nodes[id];
      if (n.op == Op::Sin) { auto& p = sc.emplace(n.a, std::make_pair(-1, -1)).first->second; p.first = id; }
      if (n.op == Op::Cos) { auto& p = sc.emplace(n.a, std::make_pair(-1, -1)).first->second; p.second = id; }
    }
    std::set<int> done_sc;
    std::ostringstream os;
    for (int id = 0; id < (int)G.nodes.size(); id++) {
      if (!live[id]) continue;
      const Node& n = G.nodes[id];
      auto A = [&] { return name(n.a); };
      auto B = [&] { return name(n.b); };
      auto fn1 = [&](const char* f) { os << "    const double t" << id << " = " << f << "(" << A() << ");\n"; };
      switch (n.op) {
        case Op::Const: case Op::Input: case Op::Param: break;
        case Op::Add: os << "    const double t" << id << " = " << A() << " + " << B() << ";\n"; break;
        case Op::Sub: os << "    const double t" << id << " = " << A() << " - " << B() << ";\n"; break;
        case Op::Mul: os << "    const double t" << id << " = " << A() << " * " << B() << ";\n"; break;
        case Op::Neg: os << "    const double t" << id << " = -" << A() << ";\n"; break;
        case Op::Recip: os << "    const double t" << id << " = hb_recip<FAST>(cx, " << A() << ");\n"; break;
        case Op::Abs: fn1("fabs"); break;
        case Op::Signum:
          os << "    const double t" << id << " = (double)((" << A() << " > 0.0) - (" << A() << " < 0.0));\n";
          break;
        case Op::Sqrt: fn1("sqrt"); break;
        case Op::Exp: os << "    const double t" << id << " = hb_exp<FAST>(cx, " << A() << ");\n"; break;
        case Op::Log: fn1("log"); break;
        case Op::Sin: case Op::Cos: {
          auto p = sc[n.a];
          if (p.first >= 0 && p.second >= 0) {
            if (!done_sc.count(n.a)) {
              done_sc.insert(n.a);
              os << "    double t" << p.first << ", t" << p.second << "; hb_sincos<FAST>(cx, " << A() << ", &t" << p.first << ", &t"
                 << p.second << ");\n";
            }
          } else {
            os << "    const double t" << id << " = " << (n.op == Op::Sin ? "hb_sin" : "hb_cos") << "<FAST>(cx, " << A() << ");\n";
          }
          break;
        }
        case Op::Tan: fn1("tan"); break;
        case Op::Asin: fn1("asin"); break;
        case Op::Acos: fn1("acos"); break;
        case Op::Atan: fn1("atan"); break;
        case Op::Sinh: fn1("sinh"); break;
        case Op::Cosh: fn1("cosh"); break;
        case Op::Tanh: fn1("tanh"); break;
        case Op::Asinh: fn1("asinh"); break;
        case Op::Acosh: fn1("acosh"); break;
        case Op::Atanh: fn1("atanh"); break;
        case Op::Pow: os << "    const double t" << id << " = pow(" << A() << ", " << B() << ");\n"; break;
        case Op::Atan2: os << "    const double t" << id << " = atan2(" << A() << ", " << B() << ");\n"; break;
      }
    }
    for (auto& o : outs) os << "    " << o.first << " = " << name(o.second) << ";\n";
    return os.str();
  }
};


// Numeric evaluation of the whole DAG (System-construction-time self-check only; the product never
// evaluates systems on the CPU).  in = [q_0..q_{n-1}, v_0..v_{n-1}].
template <class T>
std::vector<T> eval_graph(const Graph& G, const std::vector<T>& in, const std::vector<T>& prm) {
  std::vector<T> val(G.nodes.size());
  for (size_t id = 0; id < G.nodes.size(); id++) {
    const Node& n = G.nodes[id];
    const T a = n.a >= 0 && n.op != Op::Input && n.op != Op::Param ? val[n.a] : T(0);
    const T b = n.b >= 0 ? val[n.b] : T(0);
    T r = T(0);
    switch (n.op) {
      case Op::Const: r = T(n.c); break;
      case Op::Input: r = n.a < (int)in.size() ? in[n.a] : T(0); break;
      case Op::Param: r = n.a < (int)prm.size() ? prm[n.a] : T(0); break;
      case Op::Add: r = a + b; break;
      case Op::Sub: r = a - b; break;
      case Op::Mul: r = a * b; break;
      case Op::Neg: r = -a; break;
      case Op::Recip: r = T(1) / a; break;
      case Op::Abs: r = std::fabs(a); break;
      case Op::Signum: r = T((a > 0) - (a < 0)); break;
      case Op::Sqrt: r = std::sqrt(a); break;
      case Op::Exp: r = std::exp(a); break;
      case Op::Log: r = std::log(a); break;
      case Op::Sin: r = std::sin(a); break;
      case Op::Cos: r = std::cos(a); break;
      case Op::Tan: r = std::tan(a); break;
      case Op::Asin: r = std::asin(a); break;
      case Op::Acos: r = std::acos(a); break;
      case Op::Atan: r = std::atan(a); break;
      case Op::Sinh: r = std::sinh(a); break;
      case Op::Cosh: r = std::cosh(a); break;
      case Op::Tanh: r = std::tanh(a); break;
      case Op::Asinh: r = std::asinh(a); break;
      case Op::Acosh: r = std::acosh(a); break;
      case Op::Atanh: r = std::atanh(a); break;
      case Op::Pow: r = std::pow(a, b); break;
      case Op::Atan2: r = std::atan2(a, b); break;
    }
    val[id] = r;
  }
  return val;
}

// arithmetic cost (DFMA-class instructions, before FMA contraction) of the sub-DAG under `roots`, not descending into `stop`
int dag_cost(const Graph& G, const std::vector<int>& roots, const std::set<int>* stop = nullptr) {
  std::vector<char> seen(G.nodes.size(), 0);
  std::vector<int> stack(roots.begin(), roots.end());
  int cost = 0;
  while (!stack.empty()) {
    int id = stack.back(); stack.pop_back();
    if (seen[id] || (stop && stop->count(id))) continue;
    seen[id] = 1;
    const Node& n = G.nodes[id];
    if (n.op == Op::Const || n.op == Op::Input || n.op == Op::Param) continue;
    if (n.op == Op::Add || n.op == Op::Sub || n.op == Op::Mul) cost += 1;
    else if (n.op == Op::Recip) cost += 5;
    else if (n.op != Op::Neg) cost += 14;
    if (n.a >= 0) stack.push_back(n.a);
    if (n.b >= 0) stack.push_back(n.b);
  }
  return cost;
}

// Symbolic hamEqs (src/Numeric/Hamilton.hs:370-387 resolved at System-construction time):
//   A_kl    = sum_i w_i J_ik J_il                                   the mass matrix, as simplified expressions of q
//   dp_j    = 1/2 v^T (dA/dq_j) v - dU/dq_j,   dA_kl/dq_j = sum_i w_i (H_i,kj J_il + J_ik H_i,lj)
// (algebraically the reference's  p . M^-1 J^T W H_j M^-1 p - dU/dq_j  with v = M^-1 p), each coefficient a
// polynomial in the DAG's atoms reduced modulo cos^2 = 1 - sin^2.  v is solved for by the engine in between:
// `hpre` evaluates A and every v-independent value (E[]), `hpost` finishes dp from E and v.
struct SymHam {
  bool ok = false;
  std::string why;                    // when !ok
  std::vector<int> A;                 // packed lower triangle, hb_tri(j, k) order
  std::vector<int> dp;                // n nodes (functions of q-atoms and v inputs)
  std::vector<int> carried;           // v-independent nodes hpost reads from E[]
  int cost_sym = 0, cost_direct = 0;
};

SymHam build_symham(Graph& G, const SystemSpec& spec, const std::vector<SJet>& x, const SJet& U, const double* bake) {
  SymHam R;
  const int m = spec.m, n = spec.n;
  PolyForm PF(G);
  std::vector<Poly> w(m);
  for (int i = 0; i < m; i++) {
    const InertiaTerm& t = spec.inertia[i];
    w[i] = t.is_param ? (bake ? PF.constant(bake[t.param]) : PF.of(G.param(t.param))) : PF.constant(t.value);
  }
  std::vector<std::map<int, Poly>> J(m);
  std::vector<std::map<std::pair<int, int>, Poly>> H(m);
  for (int i = 0; i < m; i++) {
    for (auto& kv : x[i].g) J[i][kv.first] = PF.of(kv.second);
    for (auto& kv : x[i].h) H[i][kv.first] = PF.of(kv.second);
  }
  if (!PF.ok()) { R.why = "polynomial expansion of J/H too large"; return R; }
  auto Hget = [&](int i, int a, int b) -> const Poly* {
    auto it = H[i].find({a < b ? a : b, a < b ? b : a});
    return it == H[i].end() ? nullptr : &it->second;
  };
  auto Jget = [&](int i, int a) -> const Poly* { auto it = J[i].find(a); return it == J[i].end() ? nullptr : &it->second; };
  // mass matrix
  std::vector<std::vector<Poly>> wJ(m, std::vector<Poly>(n));
  for (int i = 0; i < m; i++)
    for (auto& kv : J[i]) wJ[i][kv.first] = PF.mul(w[i], kv.second);
  for (int j = 0; j < n; j++)
    for (int k = 0; k <= j; k++) {
      Poly acc;
      for (int i = 0; i < m; i++) {
        const Poly* a = Jget(i, k);
        if (!a || wJ[i][j].empty()) continue;
        acc = PF.add(acc, PF.mul(wJ[i][j], *a));
      }
      if (!PF.ok()) { R.why = "polynomial expansion of the mass matrix too large"; return R; }
      R.A.push_back(PF.emit(acc));
    }
  // force polynomials in the velocity inputs
  std::vector<int> vin(n);
  for (int k = 0; k < n; k++) vin[k] = G.input(n + k);
  for (int j = 0; j < n; j++) {
    Poly P;
    for (int k = 0; k < n; k++)
      for (int l = k; l < n; l++) {
        Poly g;
        for (int i = 0; i < m; i++) {
          const Poly* hk = Hget(i, k, j);
          const Poly* hl = Hget(i, l, j);
          if (hk && !wJ[i][l].empty()) g = PF.add(g, PF.mul(*hk, wJ[i][l]));
          if (hl && !wJ[i][k].empty()) g = PF.add(g, PF.mul(wJ[i][k], *hl));
        }
        if (!PF.ok()) { R.why = "polynomial expansion of dM/dq too large"; return R; }
        if (k == l) g = PF.scale(0.5, g);
        if (g.empty()) continue;
        double sign;
        const int node = PF.emit_abs(g, &sign);
        Poly term = PF.scale(sign, PF.atom(node));
        term = PF.mul(term, PF.of(vin[k]));
        term = PF.mul(term, PF.of(vin[l]));
        P = PF.add(P, term);
      }
    auto gu = U.g.find(j);
    if (gu != U.g.end() && !G.is_zero(gu->second)) {
      int node = gu->second;
      double sign = 1.0;
      if (G.nodes[node].op == Op::Neg) { node = G.nodes[node].a; sign = -1.0; }
      double c;
      if (G.is_const(node, &c)) P = PF.sub(P, PF.constant(sign * c));
      else P = PF.sub(P, PF.scale(sign, PF.atom(node)));
    }
    if (!PF.ok()) { R.why = "force polynomial too large"; return R; }
    R.dp.push_back(PF.emit(P));
  }
  // v-dependence (operands precede their users in the node list)
  std::vector<char> dep(G.nodes.size(), 0);
  for (size_t id = 0; id < G.nodes.size(); id++) {
    const Node& nd = G.nodes[id];
    if (nd.op == Op::Input) { dep[id] = nd.a >= n; continue; }
    if (nd.op == Op::Const || nd.op == Op::Param) continue;
    dep[id] = (nd.a >= 0 && dep[nd.a]) || (nd.b >= 0 && dep[nd.b]);
  }
  // Values handed from hpre to hpost.  Small systems carry every velocity-independent sub-expression (all of it is
  // then evaluated before the solve, off the critical path).  When that set is large the live state across the solve
  // (A + E) no longer fits the register file, so only the non-polynomial atoms (sin/cos/exp/recip ... results) are
  // carried and hpost re-forms the polynomial coefficients from them.
  auto cut = [&](bool atoms_only) {
    std::set<int> seen, carried;
    std::vector<int> stack(R.dp.begin(), R.dp.end());
    while (!stack.empty()) {
      int id = stack.back(); stack.pop_back();
      if (!seen.insert(id).second) continue;
      const Node& nd = G.nodes[id];
      if (nd.op == Op::Const || nd.op == Op::Input || nd.op == Op::Param) continue;
      const bool poly = nd.op == Op::Add || nd.op == Op::Sub || nd.op == Op::Mul || nd.op == Op::Neg;
      if (!dep[id] && nd.op != Op::Neg && !(atoms_only && poly)) { carried.insert(id); continue; }   // computed before the solve
      if (nd.a >= 0) stack.push_back(nd.a);
      if (nd.b >= 0) stack.push_back(nd.b);
    }
    R.carried.assign(carried.begin(), carried.end());
  };
  cut(false);
  if ((int)R.carried.size() + n * (n + 1) / 2 > 56) cut(true);
  // cost model: symbolic (hpre + hpost) vs the engine's direct sparse contraction
  {
    std::vector<int> roots = R.A;
    roots.insert(roots.end(), R.carried.begin(), R.carried.end());
    std::set<int> stop(R.carried.begin(), R.carried.end());
    R.cost_sym = dag_cost(G, roots) + dag_cost(G, R.dp, &stop);
    std::vector<int> droots;
    int nj = 0, nh = 0, prods = 0;
    std::set<std::pair<int, int>> groups;
    for (int i = 0; i < m; i++) {
      for (auto& kv : x[i].g) { droots.push_back(kv.second); nj++; }
      for (auto& kv : x[i].h) { droots.push_back(kv.second); nh++; groups.insert(kv.first); }
      for (auto& a : x[i].g) for (auto& b : x[i].g) if (a.first >= b.first) prods++;
    }
    for (auto& kv : U.g) droots.push_back(kv.second);
    int fin = 0;
    for (auto& g : groups) fin += g.first == g.second ? 1 : 2;
    R.cost_direct = dag_cost(G, droots) + nj /*W J*/ + prods /*J^T W J*/ + nj /*a = W J v*/ + nh + fin + n;
  }
  // numeric self-check: normal-form expansion must not cost accuracy (long double direct form vs double symbolic form)
  {
    unsigned long long st = 0x9E3779B97F4A7C15ULL;
    auto rnd = [&] { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) * (1.0 / 9007199254740992.0); };
    const int NS = 12;
    int valid = 0;
    const int NT = n * (n + 1) / 2;
    std::vector<long double> sa(NT, 0.0L), sd(n, 0.0L), ea(NT, 0.0L), ed(n, 0.0L);
    for (int s = 0; s < NS; s++) {
      std::vector<double> in(2 * n), prm(spec.n_params > 0 ? spec.n_params : 1);
      const double span = s < NS / 2 ? 3.0 : 0.9;
      for (int k = 0; k < n; k++) { in[k] = (2 * rnd() - 1) * span; in[n + k] = (2 * rnd() - 1) * 2.0; }
      if (s % 3 == 2) for (int k = 0; k < n; k++) in[k] = std::fabs(in[k]) + 0.25;   // radial coordinates, logs
      for (int k = 0; k < spec.n_params; k++) prm[k] = bake ? bake[k] : 0.5 + 2.0 * rnd();
      std::vector<long double> inl(in.begin(), in.end()), prml(prm.begin(), prm.end());
      const std::vector<long double> ref = eval_graph<long double>(G, inl, prml);
      const std::vector<double> val = eval_graph<double>(G, in, prm);
      std::vector<long double> wv(m);
      for (int i = 0; i < m; i++) {
        const InertiaTerm& t = spec.inertia[i];
        wv[i] = t.is_param ? prml[t.param] : (long double)t.value;
      }
      bool finite = true;
      std::vector<long double> Aref(NT, 0.0L), Aabs(NT, 0.0L), dref(n, 0.0L), dabs(n, 0.0L), a(m, 0.0L);
      for (int i = 0; i < m; i++) {
        for (auto& p : x[i].g) a[i] += wv[i] * ref[p.second] * inl[n + p.first];
        for (auto& p : x[i].g)
          for (auto& q2 : x[i].g)
            if (p.first >= q2.first) {
              const long double t = wv[i] * ref[p.second] * ref[q2.second];
              Aref[p.first * (p.first + 1) / 2 + q2.first] += t;
              Aabs[p.first * (p.first + 1) / 2 + q2.first] += fabsl(t);
            }
      }
      for (int i = 0; i < m; i++)
        for (auto& h : x[i].h) {
          const int j = h.first.first, k = h.first.second;
          const long double hv = ref[h.second];
          dref[j] += a[i] * hv * inl[n + k]; dabs[j] += fabsl(a[i] * hv * inl[n + k]);
          if (j != k) { dref[k] += a[i] * hv * inl[n + j]; dabs[k] += fabsl(a[i] * hv * inl[n + j]); }
        }
      for (auto& g : U.g) { dref[g.first] -= ref[g.second]; dabs[g.first] += fabsl(ref[g.second]); }
      for (int t = 0; t < NT; t++) finite = finite && std::isfinite((double)Aref[t]) && std::isfinite(val[R.A[t]]);
      for (int j = 0; j < n; j++) finite = finite && std::isfinite((double)dref[j]) && std::isfinite(val[R.dp[j]]);
      if (!finite) continue;
      valid++;
      for (int t = 0; t < NT; t++) { sa[t] = std::max(sa[t], Aabs[t]); ea[t] = std::max(ea[t], fabsl((long double)val[R.A[t]] - Aref[t])); }
      for (int j = 0; j < n; j++) { sd[j] = std::max(sd[j], dabs[j]); ed[j] = std::max(ed[j], fabsl((long double)val[R.dp[j]] - dref[j])); }
    }
    if (valid < 3) { R.why = "self-check found too few finite sample points"; return R; }
    for (int t = 0; t < NT; t++) if (ea[t] > 2e-14L * sa[t]) { R.why = "self-check: mass matrix entry lost accuracy in normal form"; return R; }
    for (int j = 0; j < n; j++) if (ed[j] > 2e-14L * sd[j]) { R.why = "self-check: force component lost accuracy in normal form"; return R; }
  }
  const char* force = std::getenv("HB_SYMH");
  if (force && force[0] == '0') { R.why = "disabled by HB_SYMH=0"; return R; }
  if (!(force && force[0] == '1') && R.cost_sym > R.cost_direct) { R.why = "direct sparse contraction is cheaper"; return R; }
  R.ok = true;
  return R;
}

std::string table_fn(const char* fname, const char* args, const char* index, const std::vector<int>& t) {
  std::ostringstream os;
  os << "  __host__ __device__ static constexpr int " << fname << "(" << args << ") {\n    constexpr int t["
     << (t.empty() ? 1 : t.size()) << "] = {";
  if (t.empty()) os << "0";
  for (size_t i = 0; i < t.size(); i++) os << (i ? ", " : "") << t[i];
  os << "};\n    return t[" << index << "];\n  }\n";
  return os.str();
}

}  // namespace

bool generate_system(const SystemSpec& spec, const std::string& name, GeneratedSystem& out, std::string& err) {
  const int m = spec.m, n = spec.n;
  if (n < 1 || n > HB_MAX_N || m < 1 || m > HB_MAX_M) { err = "system dimensions out of range (1 <= n <= 16, 1 <= m <= 48)"; return false; }
  if ((int)spec.inertia.size() != m || (int)spec.f_outs.size() != m) { err = "inertia / f outputs must have m entries"; return false; }
  if (spec.n_params < 0 || spec.n_params > 32) { err = "at most 32 runtime parameters"; return false; }
  for (int o : spec.f_outs) if (o < 0 || o >= (int)spec.f_ops.size()) { err = "f output index out of range"; return false; }
  if (spec.u_out < 0 || spec.u_out >= (int)spec.u_ops.size()) { err = "u output index out of range"; return false; }

  const double* bake = nullptr;
  if (!spec.baked_params.empty()) {
    if ((int)spec.baked_params.size() != spec.n_params) { err = "baked_params must have n_params entries"; return false; }
    bake = spec.baked_params.data();
  }
  Graph G;
  JetAlgebra A2(G, 2), A1(G, 1);
  std::vector<SJet> qj;
  for (int j = 0; j < n; j++) qj.push_back(A2.variable(G.input(j), j));
  std::vector<SJet> fn;
  if (!replay_tape(A2, spec.f_ops.data(), (int)spec.f_ops.size(), qj, spec.n_params, fn, err, bake)) { err = "f: " + err; return false; }
  std::vector<SJet> x;
  for (int i = 0; i < m; i++) x.push_back(fn[spec.f_outs[i]]);
  // U: first-order is all hamEqs needs (`grad u`, src/Numeric/Hamilton.hs:224); mkSystem' composes u . f (:254)
  std::vector<SJet> uin = spec.u_on_cartesian ? x : qj;
  for (auto& j : uin) j.h.clear();
  std::vector<SJet> un;
  if (!replay_tape(A1, spec.u_ops.data(), (int)spec.u_ops.size(), uin, spec.n_params, un, err, bake)) { err = "u: " + err; return false; }
  const SJet& U = un[spec.u_out];

  // structural non-zeros
  std::vector<int> jidx((size_t)m * n, -1), jrow, jcol, jnode;
  for (int i = 0; i < m; i++)
    for (auto& kv : x[i].g) { jidx[(size_t)i * n + kv.first] = (int)jrow.size(); jrow.push_back(i); jcol.push_back(kv.first); jnode.push_back(kv.second); }
  std::vector<int> hrow, hj, hk, hnode;
  for (int i = 0; i < m; i++)
    for (auto& kv : x[i].h) { hrow.push_back(i); hj.push_back(kv.first.first); hk.push_back(kv.first.second); hnode.push_back(kv.second); }
  const int NJ = (int)jrow.size(), NH = (int)hrow.size();
  // Hessian entries grouped by their (j, k) pair: dp_j gets (sum_i a_i H_i,jk) * v_k, one multiply per group
  std::vector<int> hgrp(NH, 0), gj, gk;
  {
    std::map<std::pair<int, int>, int> gid;
    for (int e = 0; e < NH; e++) {
      auto key = std::make_pair(hj[e], hk[e]);
      auto it = gid.find(key);
      if (it == gid.end()) { it = gid.emplace(key, (int)gj.size()).first; gj.push_back(hj[e]); gk.push_back(hk[e]); }
      hgrp[e] = it->second;
    }
  }
  const int NG = (int)gj.size();

  const SymHam SH = build_symham(G, spec, x, U, bake);

  Emitter E(G);
  std::ostringstream os;
  os << "// generated by hamilton_b200 sysgen: symbolic 2nd-order forward-mode derivatives of the user's tapes\n";
  os << "// hamEqs form: " << (SH.ok ? "symbolic mass matrix + force polynomials" : "direct sparse contraction (" + SH.why + ")")
     << "; cost model: symbolic " << SH.cost_sym << ", direct " << SH.cost_direct << " arithmetic ops per RHS\n";
  os << "struct " << name << " {\n";
  os << "  static constexpr int M = " << m << ", N = " << n << ", NJ = " << NJ << ", NH = " << NH << ", NG = " << NG << ", NP = " << spec.n_params << ";\n";
  os << "  static constexpr bool SYMH = " << (SH.ok ? "true" : "false") << ";\n";
  os << "  static constexpr int NE = " << (SH.ok ? (int)SH.carried.size() : 0) << ";\n";
  bool trig = false;
  for (const Node& nd : G.nodes) trig = trig || nd.op == Op::Sin || nd.op == Op::Cos || nd.op == Op::Exp;   // needs the staged table image
  os << "  static constexpr bool TRIG = " << (trig ? "true" : "false") << ";\n";
  // HEAVY: is one RK4 step of a trajectory issue-bound or HBM-bound?  Issue clocks per tile of 32 trajectories and SM:
  // (2 clk x (4 RHS x (cost model + solve) + 14 n of RK4 algebra) + 60 of bookkeeping) / 4 schedulers; HBM clocks per tile and
  // SM: 32 x 32 n bytes / 22.4 bytes per clock (6.5 TB/s over 148 SMs at 1.965 GHz).  Issue-bound systems stage their next
  // Phase with cp.async and run one CTA per SM; HBM-bound ones keep plain loads, an L2 prefetch two tiles ahead and full
  // occupancy (cp.async costs them 30-50 %: profiles/r2e).
  const double solve_cost = n * n * n / 6.0 + n * n + 5.0 * n;
  const double rhs_cost = SH.ok ? SH.cost_sym : SH.cost_direct;
  const double issue_clk = (2.0 * (4.0 * (rhs_cost + solve_cost) + 14.0 * n) + 60.0) / 4.0, hbm_clk = 32.0 * 32.0 * n / 22.4;
  const bool heavy = issue_clk >= 1.5 * hbm_clk;
  os << "  static constexpr bool HEAVY = " << (heavy ? "true" : "false") << ";   // issue " << (int)issue_clk << " vs HBM " << (int)hbm_clk << " clocks per tile\n";
  os << table_fn("jidx", "int i, int j", "i * N + j", jidx);
  os << table_fn("jrow", "int e", "e", jrow) << table_fn("jcol", "int e", "e", jcol);
  os << table_fn("hrow", "int e", "e", hrow) << table_fn("hj", "int e", "e", hj) << table_fn("hk", "int e", "e", hk);
  os << table_fn("hgrp", "int e", "e", hgrp) << table_fn("gj", "int g", "g", gj) << table_fn("gk", "int g", "g", gk);

  os << "  __device__ static __forceinline__ void inertia(const double* __restrict__ prm, double* w) {\n    (void)prm;\n";
  for (int i = 0; i < m; i++) {
    const InertiaTerm& t = spec.inertia[i];
    if (t.is_param && (t.param < 0 || t.param >= spec.n_params)) { err = "inertia parameter index out of range"; return false; }
    os << "    w[" << i << "] = " << (t.is_param ? (bake ? lit(bake[t.param]) : "prm[" + std::to_string(t.param) + "]") : lit(t.value)) << ";\n";
  }
  os << "  }\n";

  auto jouts = [&] { std::vector<std::pair<std::string, int>> o; for (int e = 0; e < NJ; e++) o.push_back({"Jv[" + std::to_string(e) + "]", jnode[e]}); return o; };
  const char* sig = "(HbCtx& cx, const double* __restrict__ prm, const double* q";
  const char* dev = "  template <bool FAST> __device__ static __forceinline__ void ";
  {
    auto o = jouts();
    for (int e = 0; e < NH; e++) o.push_back({"Hv[" + std::to_string(e) + "]", hnode[e]});
    for (int j = 0; j < n; j++) { auto it = U.g.find(j); o.push_back({"gU[" + std::to_string(j) + "]", it == U.g.end() ? G.constant(0.0) : it->second}); }
    os << "  // J non-zeros, Hessian non-zeros and grad U in one pass (shared sub-expressions)\n";
    os << dev << "derivs" << sig << ", double* Jv, double* Hv, double* gU) {\n    (void)cx; (void)prm; (void)q; (void)Jv; (void)Hv;\n" << E.body(o) << "  }\n";
  }
  os << dev << "jac" << sig << ", double* Jv) {\n    (void)cx; (void)prm; (void)q; (void)Jv;\n" << E.body(jouts()) << "  }\n";
  {
    auto o = jouts();
    o.push_back({"U", U.v});
    os << dev << "jac_pot" << sig << ", double* Jv, double& U) {\n    (void)cx; (void)prm; (void)q; (void)Jv;\n" << E.body(o) << "  }\n";
  }
  {
    std::vector<std::pair<std::string, int>> o;
    for (int i = 0; i < m; i++) o.push_back({"x[" + std::to_string(i) + "]", x[i].v});
    os << dev << "pos" << sig << ", double* x) {\n    (void)cx; (void)prm; (void)q;\n" << E.body(o) << "  }\n";
  }
  if (SH.ok) {
    const int NT = n * (n + 1) / 2;
    auto aouts = [&] { std::vector<std::pair<std::string, int>> o; for (int t = 0; t < NT; t++) o.push_back({"A[" + std::to_string(t) + "]", SH.A[t]}); return o; };
    {
      auto o = aouts();
      for (size_t e = 0; e < SH.carried.size(); e++) o.push_back({"E[" + std::to_string(e) + "]", SH.carried[e]});
      os << "  // symbolic hamEqs, part 1: packed mass matrix A(q) and every velocity-independent value of the force (E)\n";
      os << dev << "hpre" << sig << ", double* A, double* E) {\n    (void)cx; (void)prm; (void)q; (void)E;\n" << E.body(o) << "  }\n";
    }
    {
      Emitter E2(G);
      E2.n_q = n;
      for (size_t e = 0; e < SH.carried.size(); e++) E2.ext[SH.carried[e]] = "E[" + std::to_string(e) + "]";
      std::vector<std::pair<std::string, int>> o;
      for (int j = 0; j < n; j++) o.push_back({"dp[" + std::to_string(j) + "]", SH.dp[j]});
      os << "  // part 2 (after the engine solved A v = p): dp_j = 1/2 v^T (dA/dq_j) v - dU/dq_j\n";
      os << dev << "hpost" << sig << ", const double* E, const double* v, double* dp) {\n    (void)cx; (void)prm; (void)q; (void)E; (void)v;\n"
         << E2.body(o) << "  }\n";
    }
    os << dev << "smass" << sig << ", double* A) {\n    (void)cx; (void)prm; (void)q;\n" << E.body(aouts()) << "  }\n";
    {
      auto o = aouts();
      o.push_back({"U", U.v});
      os << dev << "smass_pot" << sig << ", double* A, double& U) {\n    (void)cx; (void)prm; (void)q;\n" << E.body(o) << "  }\n";
    }
  }
  os << "};\n";

  out.name = name;
  out.source = os.str();
  out.m = m; out.n = n; out.nj = NJ; out.nh = NH;
  out.n_nodes = (int)G.nodes.size();
  out.ne = SH.ok ? (int)SH.carried.size() : 0;
  out.rhs_cost = SH.ok ? SH.cost_sym : SH.cost_direct;
  out.heavy = heavy;
  out.intensity = issue_clk / hbm_clk;
  out.trig = trig;
  return true;
}

std::string jit_translation_unit(const GeneratedSystem& g, const std::string& prefix, const std::string& kind) {
  const std::string macro = kind.empty() ? "HB_DEFINE_KERNELS" : "HB_DEFINE_KERNEL_" + kind;
  // large systems: the dynamic shared memory of the RK vectors leaves room for the 8 KB sin/cos table only (HB_BIG_N = 8)
  return std::string(g.n >= 8 ? "#define HB_SC_LOG2 9\n" : "") + "#include \"hb_engine.cuh\"\n" + g.source + macro + "(" + g.name + ", " + prefix + ")\n";
}

}  // namespace hb
