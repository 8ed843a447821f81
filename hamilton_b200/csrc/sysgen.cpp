// sysgen.cpp — prints the derivative DAG of a System as a `Sys` struct for the CUDA engine.
#include "sysgen.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <sstream>

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
  explicit Emitter(const Graph& g) : G(g) {}

  std::string name(int id) const {
    const Node& n = G.nodes[id];
    switch (n.op) {
      case Op::Const: return lit(n.c);
      case Op::Input: return "q[" + std::to_string(n.a) + "]";
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
      if (live[id]) continue;
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
        case Op::Exp: fn1("exp"); break;
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

  Emitter E(G);
  std::ostringstream os;
  os << "// generated by hamilton_b200 sysgen: symbolic 2nd-order forward-mode derivatives of the user's tapes\n";
  os << "struct " << name << " {\n";
  os << "  static constexpr int M = " << m << ", N = " << n << ", NJ = " << NJ << ", NH = " << NH << ", NG = " << NG << ", NP = " << spec.n_params << ";\n";
  bool trig = false;
  for (const Node& nd : G.nodes) trig = trig || nd.op == Op::Sin || nd.op == Op::Cos;
  os << "  static constexpr bool TRIG = " << (trig ? "true" : "false") << ";\n";
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
  os << "};\n";

  out.name = name;
  out.source = os.str();
  out.m = m; out.n = n; out.nj = NJ; out.nh = NH;
  out.n_nodes = (int)G.nodes.size();
  return true;
}

std::string jit_translation_unit(const GeneratedSystem& g, const std::string& prefix, const std::string& kind) {
  const std::string macro = kind.empty() ? "HB_DEFINE_KERNELS" : "HB_DEFINE_KERNEL_" + kind;
  return "#include \"hb_engine.cuh\"\n" + g.source + macro + "(" + g.name + ", " + prefix + ")\n";
}

}  // namespace hb
