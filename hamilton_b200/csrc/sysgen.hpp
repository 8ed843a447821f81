// sysgen.hpp — System specification -> specialised CUDA source ("mkSystem as a compiler").
#pragma once
#include <string>
#include <vector>

#include "../../include/hamilton_b200.h"

namespace hb {

struct InertiaTerm {
  bool is_param = false;
  int param = 0;
  double value = 0.0;
};

// What mkSystem / mkSystem' receive (src/Numeric/Hamilton.hs:201-254), in tape form.
struct SystemSpec {
  int m = 0, n = 0, n_params = 0;
  std::vector<InertiaTerm> inertia;   // m entries
  std::vector<hb_op> f_ops;           // f : R^n -> R^m
  std::vector<int> f_outs;            // m node indices
  std::vector<hb_op> u_ops;           // u : R^n -> R  (or R^m -> R when u_on_cartesian)
  int u_out = 0;
  bool u_on_cartesian = false;
  std::vector<double> baked_params;   // if n_params values are given, PARAM leaves become literals (specialised build)
};

struct GeneratedSystem {
  std::string name;      // struct name
  std::string source;    // the struct definition (no includes, no kernels)
  int m = 0, n = 0, nj = 0, nh = 0;
  int n_nodes = 0;       // size of the derivative DAG (diagnostics)
  int ne = 0;            // values hpre hands to hpost (Sys::NE; 0 when the direct contraction is used)
  int rhs_cost = 0;      // cost model of one hamEqs evaluation in the emitted form (arithmetic ops; transcendental = 14), without the solve
  bool trig = false;     // uses sin / cos / exp (Sys::TRIG): its kernels stage the table image
  double intensity = 0;  // issue clocks / HBM clocks of one RK4 step per tile (the estimate behind `heavy`)
  bool heavy = false;    // one RK4 step costs clearly more issue time than its 32n bytes cost HBM time (Sys::HEAVY)
};

// Differentiates the tapes symbolically and prints `struct <name> { ... }` for engine/hb_engine.cuh.
bool generate_system(const SystemSpec& spec, const std::string& name, GeneratedSystem& out, std::string& err);

// Full NVRTC translation unit for one system: engine include + struct + HB_DEFINE_KERNELS(name, prefix).
// `kind` empty: all nine kernels; else only that one (lazy JIT of large systems).
std::string jit_translation_unit(const GeneratedSystem& g, const std::string& prefix, const std::string& kind = "");

}  // namespace hb
