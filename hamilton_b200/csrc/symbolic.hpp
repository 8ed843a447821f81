// symbolic.hpp — the host-side "AD at System-construction time" core.
//
// The reference differentiates the user's coordinate map with the `ad` package on every RHS call
// (jacobianT / hessianF / grad, src/Numeric/Hamilton.hs:221-224).  Here the same derivatives are
// taken ONCE, symbolically, when the System is built: the tape is replayed on second-order
// forward-mode jets whose components are nodes of a hash-consed expression DAG with algebraic
// zero/one folding, so structural zeros of J and of the Hessian tensor vanish and common
// sub-expressions (the sin/cos of each angle) are shared.  sysgen.cpp prints the DAG as straight-line
// CUDA that the hand-written engine (engine/hb_engine.cuh) inlines.
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/hamilton_b200.h"

namespace hb {

enum class Op : uint8_t {
  Const, Input, Param, Add, Sub, Mul, Neg, Recip, Abs, Signum, Sqrt, Exp, Log, Sin, Cos, Tan, Asin, Acos,
  Atan, Sinh, Cosh, Tanh, Asinh, Acosh, Atanh, Pow, Atan2
};

struct Node {
  Op op;
  int a = -1, b = -1;
  double c = 0.0;
};

class Graph {
 public:
  std::vector<Node> nodes;

  int constant(double c);
  int input(int idx) { return intern({Op::Input, idx, -1, 0.0}); }
  int param(int idx) { return intern({Op::Param, idx, -1, 0.0}); }
  int add(int a, int b);
  int sub(int a, int b);
  int mul(int a, int b);
  int neg(int a);
  int recip(int a);
  int div(int a, int b) { return mul(a, recip(b)); }
  int unary(Op op, int a);
  int pow(int a, int b);
  int atan2(int a, int b);
  int scale(double c, int a) { return mul(constant(c), a); }

  bool is_const(int id, double* v = nullptr) const {
    if (nodes[id].op != Op::Const) return false;
    if (v) *v = nodes[id].c;
    return true;
  }
  bool is_zero(int id) const { return nodes[id].op == Op::Const && nodes[id].c == 0.0; }
  bool is_one(int id) const { return nodes[id].op == Op::Const && nodes[id].c == 1.0; }

 private:
  struct Key {
    uint8_t op; int a, b; uint64_t c;
    bool operator==(const Key& o) const { return op == o.op && a == o.a && b == o.b && c == o.c; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      uint64_t h = k.op * 0x9E3779B97F4A7C15ULL;
      h ^= (uint64_t)(uint32_t)k.a * 0xBF58476D1CE4E5B9ULL + (h << 6) + (h >> 2);
      h ^= (uint64_t)(uint32_t)k.b * 0x94D049BB133111EBULL + (h << 6) + (h >> 2);
      h ^= k.c + (h << 6) + (h >> 2);
      return (size_t)h;
    }
  };
  std::unordered_map<Key, int, KeyHash> cse_;
  int intern(const Node& n);
};

// Second-order forward-mode jet over n independent variables with sparse, symbolic components.
struct SJet {
  int v = -1;                                // value node
  std::map<int, int> g;                      // d/dq_j          (absent = structurally zero)
  std::map<std::pair<int, int>, int> h;      // d2/dq_j dq_k, j <= k
};

class JetAlgebra {
 public:
  JetAlgebra(Graph& g, int order) : G(g), order_(order) {}
  Graph& G;
  SJet constant(double c) { SJet r; r.v = G.constant(c); return r; }
  SJet leaf(int node) { SJet r; r.v = node; return r; }
  SJet variable(int node, int j) { SJet r; r.v = node; r.g[j] = G.constant(1.0); return r; }
  SJet add(const SJet& a, const SJet& b);
  SJet sub(const SJet& a, const SJet& b);
  SJet mul(const SJet& a, const SJet& b);
  SJet neg(const SJet& a);
  SJet recip(const SJet& a);
  SJet div(const SJet& a, const SJet& b) { return mul(a, recip(b)); }
  SJet unary(int hb_opcode, const SJet& a);
  SJet powi(const SJet& a, int k);
  SJet pow(const SJet& a, const SJet& b);
  SJet atan2(const SJet& a, const SJet& b);
  // r = phi(a) given nodes for phi(a.v), phi'(a.v), phi''(a.v)
  SJet chain(const SJet& a, int f0, int f1, int f2);

 private:
  int order_;
  void acc(std::map<int, int>& m, int key, int node);
  void acc(std::map<std::pair<int, int>, int>& m, std::pair<int, int> key, int node);
};

// Replays a tape on jets.  `inputs` are the jets of the tape's inputs.  Returns false with a message
// on a malformed tape (forward reference, bad opcode ...).
// `bake` (optional, n_params values) turns PARAM leaves into literals so they fold like constants.
bool replay_tape(JetAlgebra& A, const hb_op* ops, int n_ops, const std::vector<SJet>& inputs, int n_params,
                 std::vector<SJet>& nodes, std::string& err, const double* bake = nullptr);

}  // namespace hb
