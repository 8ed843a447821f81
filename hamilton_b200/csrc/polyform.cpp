// polyform.cpp — polynomial normal form with Pythagorean reduction and greedy Horner emission.
#include "polyform.hpp"

#include <algorithm>
#include <cmath>

namespace hb {

void PolyForm::acc(Poly& p, const Mono& m, double c) {
  if (c == 0.0) return;
  auto it = p.find(m);
  if (it == p.end()) { p.emplace(m, c); return; }
  const double big = std::max(std::fabs(it->second), std::fabs(c));
  it->second += c;
  // literal coefficients that were multiplied in different orders cancel only to within a few ulp
  if (std::fabs(it->second) <= 64 * 2.220446049250313e-16 * big) p.erase(it);
}

Mono PolyForm::mono_mul(const Mono& a, const Mono& b) {
  Mono r;
  size_t i = 0, j = 0;
  while (i < a.f.size() || j < b.f.size()) {
    if (j == b.f.size() || (i < a.f.size() && a.f[i].first < b.f[j].first)) r.f.push_back(a.f[i++]);
    else if (i == a.f.size() || b.f[j].first < a.f[i].first) r.f.push_back(b.f[j++]);
    else {
      const int pw = a.f[i].second + b.f[j].second;
      if (pw != 0) r.f.push_back({a.f[i].first, pw});
      i++; j++;
    }
  }
  return r;
}

Poly PolyForm::constant(double c) const {
  Poly p;
  if (c != 0.0) p.emplace(Mono(), c);
  return p;
}

Poly PolyForm::atom(int node) {
  opaque_[node] = 1;
  Poly p;
  Mono m;
  m.f.push_back({node, 1});
  p.emplace(m, 1.0);
  return p;
}

Poly PolyForm::add(const Poly& a, const Poly& b) const {
  Poly r = a;
  for (auto& kv : b) acc(r, kv.first, kv.second);
  return r;
}
Poly PolyForm::sub(const Poly& a, const Poly& b) const {
  Poly r = a;
  for (auto& kv : b) acc(r, kv.first, -kv.second);
  return r;
}
Poly PolyForm::scale(double c, const Poly& a) const {
  Poly r;
  if (c == 0.0) return r;
  for (auto& kv : a) r.emplace(kv.first, c * kv.second);
  return r;
}

Poly PolyForm::mul(const Poly& a, const Poly& b) {
  Poly r;
  if (!ok_) return r;
  if (a.size() * b.size() > 64 * max_terms_) { ok_ = false; return r; }
  for (auto& x : a)
    for (auto& y : b) acc(r, mono_mul(x.first, y.first), x.second * y.second);
  reduce_trig(r);
  if (r.size() > max_terms_) { ok_ = false; r.clear(); }
  return r;
}

// cos(a)^p -> cos(a)^(p mod 2) * (1 - sin(a)^2)^(p div 2): cos degree <= 1 is a canonical form for polynomials
// in (sin a, cos a), so everything the Pythagorean identity can cancel does cancel.
void PolyForm::reduce_trig(Poly& p) {
  if (cos_to_sin_.empty()) return;
  for (int guard = 0; guard < 64; guard++) {
    bool changed = false;
    Poly out;
    for (auto& kv : p) {
      const Mono& m = kv.first;
      int at = -1;
      for (size_t i = 0; i < m.f.size(); i++)
        if (m.f[i].second >= 2 && cos_to_sin_.count(m.f[i].first)) { at = (int)i; break; }
      if (at < 0) { acc(out, m, kv.second); continue; }
      changed = true;
      const int c = m.f[at].first, s = cos_to_sin_[c], pw = m.f[at].second, h = pw / 2;
      Mono rest;
      for (size_t i = 0; i < m.f.size(); i++)
        if ((int)i != at) rest.f.push_back(m.f[i]);
      if (pw & 1) { Mono one; one.f.push_back({c, 1}); rest = mono_mul(rest, one); }
      double binom = 1.0;   // C(h, t) (-1)^t
      for (int t = 0; t <= h; t++) {
        Mono sp;
        if (t > 0) sp.f.push_back({s, 2 * t});
        acc(out, mono_mul(rest, sp), kv.second * binom);
        binom = -binom * (double)(h - t) / (double)(t + 1);
      }
    }
    p.swap(out);
    if (!changed) break;
  }
}

Poly PolyForm::of(int node) {
  auto it = memo_.find(node);
  if (it != memo_.end()) return it->second;
  Poly r;
  if (opaque_.count(node)) { r = atom(node); memo_[node] = r; return r; }
  const Node n = G.nodes[node];
  switch (n.op) {
    case Op::Const: r = constant(n.c); break;
    case Op::Add: r = add(of(n.a), of(n.b)); break;
    case Op::Sub: r = sub(of(n.a), of(n.b)); break;
    case Op::Neg: r = scale(-1.0, of(n.a)); break;
    case Op::Mul: r = mul(of(n.a), of(n.b)); break;
    case Op::Recip: {
      Poly pa = of(n.a);
      if (pa.size() == 1) {   // 1 / (c * monomial): negative powers, so x * (1/x) cancels
        Mono inv = pa.begin()->first;
        for (auto& f : inv.f) f.second = -f.second;
        r.emplace(inv, 1.0 / pa.begin()->second);
      } else {
        r = atom(node);
      }
      break;
    }
    case Op::Pow: {
      double e;
      if (G.is_const(n.b, &e) && e == std::floor(e) && e >= 1 && e <= 8) {
        Poly base = of(n.a);
        r = constant(1.0);
        for (int k = 0; k < (int)e; k++) r = mul(r, base);
      } else {
        r = atom(node);
      }
      break;
    }
    case Op::Cos:
      cos_to_sin_[node] = G.unary(Op::Sin, n.a);
      r = atom(node);
      break;
    default: r = atom(node);
  }
  memo_[node] = r;
  return r;
}

// ------------------------------------------------------------------------------ emission ----
int PolyForm::emit_pow(int a, int k) {
  int r = -1, base = a;
  while (k) {
    if (k & 1) r = r < 0 ? base : G.mul(r, base);
    k >>= 1;
    if (k) base = G.mul(base, base);
  }
  return r;
}

int PolyForm::emit_mono(double c, const Mono& m) {
  int x = -1;
  for (auto& f : m.f)
    if (f.second > 0) { int t = emit_pow(f.first, f.second); x = x < 0 ? t : G.mul(x, t); }
  for (auto& f : m.f)
    if (f.second < 0) { int t = emit_pow(G.recip(f.first), -f.second); x = x < 0 ? t : G.mul(x, t); }
  if (x < 0) return G.constant(c);
  return G.mul(G.constant(c), x);
}

int PolyForm::emit(const Poly& p) {
  if (p.empty()) return G.constant(0.0);
  double sign;
  const int n = emit_abs(p, &sign);
  return sign < 0 ? G.neg(n) : n;
}

int PolyForm::emit_abs(const Poly& p, double* sign) {
  if (p.empty()) { *sign = 1.0; return G.constant(0.0); }
  const bool negate = p.begin()->second < 0;
  *sign = negate ? -1.0 : 1.0;
  const Poly q = negate ? scale(-1.0, p) : p;
  auto it = emitted_.find(q);
  if (it != emitted_.end()) return it->second;
  const int node = emit_rec(q);
  emitted_[q] = node;
  return node;
}

static int mono_degree(const Mono& m) {
  int d = 0;
  for (auto& f : m.f) d += f.second < 0 ? -f.second : f.second;
  return d;
}

int PolyForm::emit_rec(const Poly& q) {
  if (q.size() == 1) return emit_mono(q.begin()->second, q.begin()->first);
  // (1) leading-coefficient normalisation, when it removes more coefficient multiplies than the one it adds
  {
    const double c0 = q.begin()->second;
    if (c0 != 1.0) {
      int costly = 0, costly_n = 1;
      for (auto& kv : q) {
        if (mono_degree(kv.first) < 2) continue;
        if (std::fabs(kv.second) != 1.0) costly++;
        if (std::fabs(kv.second / c0) != 1.0) costly_n++;
      }
      if (costly_n < costly) return G.mul(G.constant(c0), emit(scale(1.0 / c0, q)));
    }
  }
  // (2) common monomial content
  {
    Mono content;
    for (auto& f : q.begin()->first.f) {
      int best = f.second;
      for (auto& kv : q) {
        int pw = 0;
        for (auto& g : kv.first.f) if (g.first == f.first) pw = g.second;
        if (f.second > 0) best = std::min(best, std::max(pw, 0)); else best = std::max(best, std::min(pw, 0));
      }
      if (best != 0) content.f.push_back({f.first, best});
    }
    if (!content.f.empty()) {
      Mono inv = content;
      for (auto& f : inv.f) f.second = -f.second;
      Poly rest;
      for (auto& kv : q) acc(rest, mono_mul(kv.first, inv), kv.second);
      return G.mul(emit_mono(1.0, content), emit(rest));
    }
  }
  // (3) greedy Horner: split on the atom shared by the most monomials
  {
    std::map<int, int> count;
    for (auto& kv : q)
      for (auto& f : kv.first.f) if (f.second > 0) count[f.first]++;
    int best = -1, best_n = 1;
    for (auto& kv : count) if (kv.second > best_n) { best = kv.first; best_n = kv.second; }
    if (best >= 0) {
      int kmin = 1 << 30;
      for (auto& kv : q)
        for (auto& f : kv.first.f) if (f.first == best && f.second > 0) kmin = std::min(kmin, f.second);
      Mono inv;
      inv.f.push_back({best, -kmin});
      Poly with, without;
      for (auto& kv : q) {
        bool has = false;
        for (auto& f : kv.first.f) if (f.first == best && f.second > 0) has = true;
        if (has) acc(with, mono_mul(kv.first, inv), kv.second); else acc(without, kv.first, kv.second);
      }
      const int head = G.mul(emit_pow(best, kmin), emit(with));
      return without.empty() ? head : G.add(head, emit(without));
    }
  }
  // (4) plain sum of monomials
  int accn = -1;
  for (auto& kv : q) {
    const int t = emit_mono(kv.second, kv.first);
    accn = accn < 0 ? t : G.add(accn, t);
  }
  return accn;
}

}  // namespace hb
