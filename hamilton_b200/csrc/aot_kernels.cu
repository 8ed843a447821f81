// aot_kernels.cu — system-independent AOT pieces: the init_random kernel and the table of the
// per-system kernel tables (gen/aot_<name>.cu, one translation unit per built-in system).
#include "../../include/hamilton_b200.h"
#include "engine/hb_engine.cuh"

extern "C" __global__ void __launch_bounds__(HB_BLOCK) hbk_init_random(const __grid_constant__ HbKArgs a) { hb_body_init_random(a); }

#define HB_DECL(name) extern "C" const void* const hb_aot_table_##name[HB_K_COUNT + 1]; extern "C" const void* const hb_aot_table_##name##_dflt[HB_K_COUNT + 1];
HB_DECL(pendulum) HB_DECL(double_pendulum) HB_DECL(room) HB_DECL(two_body) HB_DECL(spring) HB_DECL(bezier)
HB_DECL(triple_pendulum) HB_DECL(chain12) HB_DECL(spring1d)

// builtin in [0, HB_SYS__COUNT): runtime-parameter kernels; builtin + HB_SYS__COUNT: default parameters baked in.
static const void* const* hb_aot_tables(int builtin);
extern "C" const void* hb_aot_kernel(int builtin, int kernel_id) {
  const void* const* t = hb_aot_tables(builtin);
  if (!t || kernel_id < 0 || kernel_id >= HB_K_COUNT) return nullptr;
  return t[kernel_id];
}
// doubles of dynamic shared memory per thread that every kernel of this system is launched with
extern "C" int hb_aot_dyn_doubles(int builtin) {
  const void* const* t = hb_aot_tables(builtin);
  return t ? (int)(unsigned long long)t[HB_K_COUNT] : 0;
}
static const void* const* hb_aot_tables(int builtin) {
  static const void* const* const tables[2 * HB_SYS__COUNT] = {   // order = hb_builtin
      hb_aot_table_pendulum, hb_aot_table_double_pendulum, hb_aot_table_room, hb_aot_table_two_body, hb_aot_table_spring,
      hb_aot_table_bezier, hb_aot_table_triple_pendulum, hb_aot_table_chain12, hb_aot_table_spring1d,
      hb_aot_table_pendulum_dflt, hb_aot_table_double_pendulum_dflt, hb_aot_table_room_dflt, hb_aot_table_two_body_dflt,
      hb_aot_table_spring_dflt, hb_aot_table_bezier_dflt, hb_aot_table_triple_pendulum_dflt, hb_aot_table_chain12_dflt,
      hb_aot_table_spring1d_dflt};
  if (builtin < 0 || builtin >= 2 * HB_SYS__COUNT) return nullptr;
  return tables[builtin];
}
extern "C" const void* hb_aot_init_random(void) { return (const void*)hbk_init_random; }
extern "C" size_t hb_aot_kargs_size(void) { return sizeof(HbKArgs); }
