// hamilton.hpp — header-only C++ mirror of Numeric.Hamilton over the C ABI (include/hamilton_b200.h).
//
// The reference is compiled Haskell; GHC is not available in this image, so this is the host-side
// mirror in the image's compiled language.  Same names, argument order and error behaviour as the
// Haskell module (src/Numeric/Hamilton.hs:28-70):
//
//   System<M, N>  mkSystem / mkSystem'      (mkSystem_)          :201-254
//   Config<N>, Phase<N>                                           :103-143
//   underlyingPos, pe                                             :174-186
//   momenta, toPhase, keC, lagrangian                             :262-309
//   velocities, fromPhase, keP, hamiltonian                       :316-361
//   hamEqs                                                        :370-387
//   stepHam, evolveHam, evolveHam' (evolveHam_)                   :390-462
//   stepHamC, evolveHamC, evolveHamC' (evolveHamC_)               :470-515
//
// mkSystem's rank-2 argument `forall a. RealFloat a => Vector n a -> Vector m a` becomes a generic
// lambda `[](auto const& q) { ... return std::array<decltype(q[0] + 0.0), M>{...}; }`, instantiated at
// the tracing type hb::Ex to record the tape that hb_system_from_tape compiles to sm_100a code.
// Failures map to exceptions exactly where the reference calls `error` / throws.
#pragma once
#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/hamilton_b200.h"
#include "../include/hamilton_b200_trace.hpp"

namespace hamilton {

template <int N> using R = std::array<double, N>;

template <int N>
struct Config {   // Cfg { cfgPositions, cfgVelocities }
  R<N> cfgPositions, cfgVelocities;
};
template <int N>
struct Phase {    // Phs { phsPositions, phsMomenta }
  R<N> phsPositions, phsMomenta;
};

struct HamiltonError : std::runtime_error {
  int status;
  HamiltonError(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};
inline void check(hb_status s) {
  if (s != HB_OK) throw HamiltonError(s, std::string("hamilton_b200 status ") + std::to_string(s) + ": " + hb_last_error());
}

template <int M, int N>
class System {   // opaque, like the reference's (constructor not exported, src/Numeric/Hamilton.hs:32)
 public:
  explicit System(hb_system* h) : h_(h, hb_system_free) {}
  hb_system* get() const { return h_.get(); }
 private:
  std::shared_ptr<hb_system> h_;
};

namespace detail {
template <int M, int N, class F, class U>
System<M, N> make(const R<M>& inertia, F&& f, U&& u, bool u_on_cartesian) {
  hb::Tape tf, tu;
  std::array<hb::Ex, N> q;
  for (int j = 0; j < N; j++) q[j] = hb::Ex::input(tf, j);
  auto x = f(q);
  static_assert(std::tuple_size<decltype(x)>::value == M, "f must return M coordinates");
  std::array<int32_t, M> fouts;
  for (int i = 0; i < M; i++) fouts[i] = x[i].id;
  int32_t uout;
  if (u_on_cartesian) {
    std::array<hb::Ex, M> xin;
    for (int i = 0; i < M; i++) xin[i] = hb::Ex::input(tu, i);
    uout = u(xin).id;
  } else {
    std::array<hb::Ex, N> qin;
    for (int j = 0; j < N; j++) qin[j] = hb::Ex::input(tu, j);
    // generic lambdas are instantiated separately for array<Ex, N> and array<Ex, M>; reuse u on q
    uout = u(qin).id;
  }
  hb_tape ft{N, (int32_t)tf.ops.size(), tf.ops.data(), M, fouts.data()};
  hb_tape ut{u_on_cartesian ? M : N, (int32_t)tu.ops.size(), tu.ops.data(), 1, &uout};
  hb_system* h = nullptr;
  check(hb_system_from_tape(M, N, inertia.data(), &ft, &ut, u_on_cartesian ? 1 : 0, nullptr, 0, &h));
  return System<M, N>(h);
}
}  // namespace detail

// mkSystem: potential on the generalized coordinates (src/Numeric/Hamilton.hs:201-225)
template <int M, int N, class F, class U>
System<M, N> mkSystem(const R<M>& inertia, F&& f, U&& u) { return detail::make<M, N>(inertia, f, u, false); }
// mkSystem': potential on the underlying Cartesian coordinates (:238-254)
template <int M, int N, class F, class U>
System<M, N> mkSystem_(const R<M>& inertia, F&& f, U&& u) { return detail::make<M, N>(inertia, f, u, true); }

template <int M, int N> R<M> underlyingPos(const System<M, N>& s, const R<N>& q) { R<M> x; check(hb_underlying_pos(s.get(), q.data(), x.data())); return x; }
template <int M, int N> double pe(const System<M, N>& s, const R<N>& q) { double u; check(hb_pe(s.get(), q.data(), &u)); return u; }
template <int M, int N> R<N> momenta(const System<M, N>& s, const Config<N>& c) { R<N> p; check(hb_momenta(s.get(), c.cfgPositions.data(), c.cfgVelocities.data(), p.data())); return p; }
template <int M, int N> Phase<N> toPhase(const System<M, N>& s, const Config<N>& c) { return Phase<N>{c.cfgPositions, momenta(s, c)}; }
template <int M, int N> R<N> velocities(const System<M, N>& s, const Phase<N>& p) { R<N> v; check(hb_velocities(s.get(), p.phsPositions.data(), p.phsMomenta.data(), v.data())); return v; }
template <int M, int N> Config<N> fromPhase(const System<M, N>& s, const Phase<N>& p) { return Config<N>{p.phsPositions, velocities(s, p)}; }
template <int M, int N> double keC(const System<M, N>& s, const Config<N>& c) { double t; check(hb_ke_c(s.get(), c.cfgPositions.data(), c.cfgVelocities.data(), &t)); return t; }
template <int M, int N> double lagrangian(const System<M, N>& s, const Config<N>& c) { double l; check(hb_lagrangian(s.get(), c.cfgPositions.data(), c.cfgVelocities.data(), &l)); return l; }
template <int M, int N> double keP(const System<M, N>& s, const Phase<N>& p) { double t; check(hb_ke_p(s.get(), p.phsPositions.data(), p.phsMomenta.data(), &t)); return t; }
template <int M, int N> double hamiltonian(const System<M, N>& s, const Phase<N>& p) { double h; check(hb_hamiltonian(s.get(), p.phsPositions.data(), p.phsMomenta.data(), &h)); return h; }

// hamEqs :: System m n -> Phase n -> (R n, R n)   (:370-387)
template <int M, int N>
std::pair<R<N>, R<N>> hamEqs(const System<M, N>& s, const Phase<N>& p) {
  R<N> dq, dp;
  check(hb_ham_eqs(s.get(), p.phsPositions.data(), p.phsMomenta.data(), dq.data(), dp.data()));
  return {dq, dp};
}
// stepHam :: Double -> System m n -> Phase n -> Phase n   (:390-402) — the reference's adaptive RKF45 solve over (0, r)
template <int M, int N>
Phase<N> stepHam(double r, const System<M, N>& s, const Phase<N>& p) {
  Phase<N> o;
  check(hb_step_ham(s.get(), r, p.phsPositions.data(), p.phsMomenta.data(), o.phsPositions.data(), o.phsMomenta.data()));
  return o;
}
// evolveHam :: System m n -> Phase n -> Vector s Double -> Vector s (Phase n), 2 <= s   (:433-462)
template <int M, int N>
std::vector<Phase<N>> evolveHam(const System<M, N>& s, const Phase<N>& p0, const std::vector<double>& ts) {
  if (ts.size() < 2) throw std::invalid_argument("evolveHam: need at least two times (2 <= s)");
  std::vector<double> out(ts.size() * 2 * N);
  check(hb_evolve_ham(s.get(), p0.phsPositions.data(), p0.phsMomenta.data(), ts.data(), (int32_t)ts.size(), out.data()));
  std::vector<Phase<N>> r(ts.size());
  for (size_t k = 0; k < ts.size(); k++)
    for (int j = 0; j < N; j++) { r[k].phsPositions[j] = out[k * 2 * N + j]; r[k].phsMomenta[j] = out[k * 2 * N + N + j]; }
  return r;
}
// evolveHam' (:409-429): [] -> []; [x] -> grid [0, x], first point dropped
template <int M, int N>
std::vector<Phase<N>> evolveHam_(const System<M, N>& s, const Phase<N>& p0, const std::vector<double>& ts) {
  if (ts.empty()) return {};
  if (ts.size() == 1) { auto r = evolveHam(s, p0, {0.0, ts[0]}); r.erase(r.begin()); return r; }
  return evolveHam(s, p0, ts);
}
template <int M, int N>
Config<N> stepHamC(double r, const System<M, N>& s, const Config<N>& c) {   // :505-515
  Config<N> o;
  check(hb_step_ham_c(s.get(), r, c.cfgPositions.data(), c.cfgVelocities.data(), o.cfgPositions.data(), o.cfgVelocities.data()));
  return o;
}
template <int M, int N>
std::vector<Config<N>> evolveHamC(const System<M, N>& s, const Config<N>& c0, const std::vector<double>& ts) {   // :488-498
  if (ts.size() < 2) throw std::invalid_argument("evolveHamC: need at least two times (2 <= s)");
  std::vector<double> out(ts.size() * 2 * N);
  check(hb_evolve_ham_c(s.get(), c0.cfgPositions.data(), c0.cfgVelocities.data(), ts.data(), (int32_t)ts.size(), out.data()));
  std::vector<Config<N>> r(ts.size());
  for (size_t k = 0; k < ts.size(); k++)
    for (int j = 0; j < N; j++) { r[k].cfgPositions[j] = out[k * 2 * N + j]; r[k].cfgVelocities[j] = out[k * 2 * N + N + j]; }
  return r;
}
template <int M, int N>
std::vector<Config<N>> evolveHamC_(const System<M, N>& s, const Config<N>& c0, const std::vector<double>& ts) {   // :470-480
  if (ts.empty()) return {};
  if (ts.size() == 1) { auto r = evolveHamC(s, c0, {0.0, ts[0]}); r.erase(r.begin()); return r; }
  return evolveHamC(s, c0, ts);
}

}  // namespace hamilton
