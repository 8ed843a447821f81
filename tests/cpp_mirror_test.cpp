// Exercises cpp/hamilton.hpp the way README.md:88-165 of the reference uses Numeric.Hamilton:
// build the double pendulum with mkSystem', convert a Config to a Phase, evolve it.  Prints a
// few numbers for the pytest driver to compare against the oracle.
#include <cstdio>
#include <cmath>
#include "../cpp/hamilton.hpp"
using namespace hamilton;
using hb::sin; using hb::cos;   // the polymorphic math of the tracing number type
int main(int argc, char** argv) {
  const bool create_only = argc > 1;
  const double m1 = 1.0, m2 = 2.0;   // README.md:92-103 variant: g = 5, masses (1, 2)
  try {
    auto dp = mkSystem_<4, 2>(R<4>{m1, m1, m2, m2},
        [](auto const& q) { return std::array<hb::Ex, 4>{sin(q[0]), -cos(q[0]), sin(q[0]) + sin(q[1]) / 2.0, -cos(q[0]) - cos(q[1]) / 2.0}; },
        [=](auto const& x) { return 5.0 * (m1 * x[1] + m2 * x[3]); });
    std::printf("created\n");
    if (create_only) return 0;
    Config<2> c0{{1.0, 0.0}, {0.0, 0.5}};                 // README.md:124-127
    Phase<2> p0 = toPhase(dp, c0);
    std::printf("momenta %.17g %.17g\n", p0.phsMomenta[0], p0.phsMomenta[1]);
    std::printf("hamiltonian %.17g\n", hamiltonian(dp, p0));
    auto de = hamEqs(dp, p0);
    std::printf("hamEqs %.17g %.17g %.17g %.17g\n", de.first[0], de.first[1], de.second[0], de.second[1]);
    std::vector<double> ts;
    for (int k = 0; k <= 10; k++) ts.push_back(0.1 * k);   // README.md:141  evolveHam doublePendulum phase0 [0,0.1 .. 1]
    auto ev = evolveHam(dp, p0, ts);
    std::printf("evolve_last %.17g %.17g %.17g %.17g\n", ev.back().phsPositions[0], ev.back().phsPositions[1], ev.back().phsMomenta[0], ev.back().phsMomenta[1]);
    Phase<2> p1 = stepHam(0.1, dp, p0);
    std::printf("step %.17g %.17g %.17g %.17g\n", p1.phsPositions[0], p1.phsPositions[1], p1.phsMomenta[0], p1.phsMomenta[1]);
    Config<2> c1 = stepHamC(0.1, dp, c0);
    std::printf("stepC %.17g %.17g %.17g %.17g\n", c1.cfgPositions[0], c1.cfgPositions[1], c1.cfgVelocities[0], c1.cfgVelocities[1]);
    std::printf("evolve_prime_sizes %zu %zu\n", evolveHam_(dp, p0, {}).size(), evolveHam_(dp, p0, {0.1}).size());
    try { evolveHam(dp, p0, {0.0}); std::printf("no_throw\n"); } catch (const std::invalid_argument&) { std::printf("throws_on_short_grid\n"); }
  } catch (const HamiltonError& e) {
    std::printf("error %d %s\n", e.status, e.what());
    return 3;
  }
  return 0;
}
