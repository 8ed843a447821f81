// polyform.hpp — sparse multivariate polynomial normal form over the atoms of an expression DAG.
//
// The engine's hot loop is bound by the FP64 pipe, so every DFMA the system compiler can prove
// unnecessary is throughput.  The reference forms M = J^T W J and the force term
// p . M^-1 J^T W H_j M^-1 p numerically on every hamEqs call (src/Numeric/Hamilton.hs:377-387); here
// both are formed ONCE, symbolically, as polynomials in the DAG's non-polynomial nodes ("atoms":
// inputs, parameters, sin/cos/exp/... nodes), reduced modulo cos^2 = 1 - sin^2 so that the
// Pythagorean cancellations every linkage has (c^2 + s^2 -> 1, -s c + c s -> 0) happen at
// System-construction time, and printed back as a greedily factored (Horner-like) DAG.
#pragma once
#include <map>
#include <unordered_map>
#include <utility>
#include <vector>

#include "symbolic.hpp"

namespace hb {

// product of atom^power, atoms sorted by node id, powers non-zero (negative = reciprocal of the atom)
struct Mono {
  std::vector<std::pair<int, int>> f;
  bool operator<(const Mono& o) const { return f < o.f; }
  bool operator==(const Mono& o) const { return f == o.f; }
};
typedef std::map<Mono, double> Poly;

class PolyForm {
 public:
  explicit PolyForm(Graph& g, size_t max_terms = 4096) : G(g), max_terms_(max_terms) {}
  Graph& G;

  bool ok() const { return ok_; }          // false once an expansion exceeded max_terms (results are then meaningless)
  Poly constant(double c) const;
  Poly atom(int node);                      // the monomial `node` (marks it opaque: never expanded)
  Poly of(int node);                        // normal form of a DAG node (memoised)
  Poly add(const Poly& a, const Poly& b) const;
  Poly sub(const Poly& a, const Poly& b) const;
  Poly mul(const Poly& a, const Poly& b);
  Poly scale(double c, const Poly& a) const;
  static bool is_zero(const Poly& p) { return p.empty(); }

  // Normal form -> DAG node.  Equal polynomials give the same node; p and -p share everything but a negation.
  int emit(const Poly& p);
  // sign-normalised emit: p = sign * node, sign = +1 / -1 chosen so that p and -p return the same node
  int emit_abs(const Poly& p, double* sign);

 private:
  size_t max_terms_;
  bool ok_ = true;
  std::unordered_map<int, Poly> memo_;
  std::map<int, int> cos_to_sin_;           // Cos(a) node -> Sin(a) node
  std::map<int, char> opaque_;
  std::map<Poly, int> emitted_;

  void reduce_trig(Poly& p);
  static void acc(Poly& p, const Mono& m, double c);
  static Mono mono_mul(const Mono& a, const Mono& b);
  int emit_mono(double c, const Mono& m);
  int emit_pow(int atom, int k);
  int emit_rec(const Poly& p);
};

}  // namespace hb
