// Entry point of a user-system plugin: `dfb_plugin_describe` hands libdiffurch_b200.so the
// static facts of the user's device functor and the launchers instantiated for it from the same
// kernel templates as the built-in systems (kernels_*.cu compiled with -DDFB_PLUGIN_SYSTEM).
// Built by diffurch_b200.user.build_user_system; loaded by dfb_register_plugin.
#ifndef DFB_PLUGIN_SYSTEM
#error "compile with -DDFB_PLUGIN_HEADER=\"...\" -DDFB_PLUGIN_SYSTEM=<struct name>"
#endif
#define DFB_ARITH 1
#include "dev/systems.cuh"  // pulls the user header in (namespace dfb_fast here)
#include "launch.hpp"
#include "launch_table.hpp"

DFB_DECLARE_SHIMS(dfb_strict)
DFB_DECLARE_SHIMS(dfb_fast)
namespace dfb_fast {
int run_dde_fixed_shim(int, const dfb_tableau&, double, double, double, double, int, int, size_t, const double*,
                       const double*, double*, double*, double*, unsigned long long*, unsigned long long*,
                       uint32_t*, int, int, double*, double*, uint32_t*, cudaStream_t);
int run_dde_event_shim(int, const dfb_tableau&, double, double, double, double, long long, int, int, unsigned,
                       double*, size_t, const double*, const double*, double*, double*, double*,
                       unsigned long long*, unsigned long long*, unsigned long long*, unsigned long long*,
                       uint32_t*, int, const dfb_locator*, int, const double*, const int32_t*, int, double*,
                       int32_t*, uint32_t*, int, int, double*, double*, uint32_t*, unsigned long long*, int, int,
                       int, int, cudaStream_t);
int dde_event_query(int, int, int);
int run_ode_adaptive_shim(int, const dfb_tableau&, const dfb_auto_stepsize&, double, double, size_t,
                          const double*, const double*, double*, double*, unsigned long long*,
                          unsigned long long*, unsigned long long*, uint32_t*, unsigned long long*, int,
                          int, int, double*, double*, uint32_t*, cudaStream_t);
}

#define DFB_STR2(x) #x
#define DFB_STR(x) DFB_STR2(x)

extern "C" __attribute__((visibility("default"))) int dfb_plugin_describe(dfb_plugin_desc* out) {
  using Sys = dfb_fast::DFB_PLUGIN_SYSTEM;
  static_assert(Sys::N >= 1 && Sys::N <= DFB_MAX_STATE, "N out of range");
  static_assert(Sys::NP <= DFB_MAX_PARAMS && Sys::NA <= DFB_MAX_AUX, "NP / NA out of range");
  if (!out) return -1;
  out->plugin_abi = DFB_PLUGIN_ABI;
  out->n_state = Sys::N;
  out->n_params = Sys::NP;
  out->n_aux = Sys::NA;
  out->n_detectors = Sys::ND;
  out->n_callbacks = Sys::NC;
  out->uses_history = Sys::uses_history ? 1 : 0;
  out->dde_hot_ok = (Sys::uses_history && Sys::has_const_delay) ? 1 : 0;
  out->sizeof_devproblem = unsigned(sizeof(DevProblem));
  out->sizeof_locator = unsigned(sizeof(dfb_locator));
  out->sizeof_tableau = unsigned(sizeof(dfb_tableau));
  out->sizeof_auto_stepsize = unsigned(sizeof(dfb_auto_stepsize));
  out->name = DFB_STR(DFB_PLUGIN_SYSTEM);
  out->fast = DfbLaunchTable{dfb_fast::run_ode_fixed_shim, dfb_fast::run_ode_fixed_periodic_shim,
                             dfb_fast::run_ode_event_shim, dfb_fast::run_ode_adaptive_shim,
                             dfb_fast::run_dde_fixed_shim, {dfb_fast::launch_general_g1}, 1,
                             dfb_fast::run_dde_event_shim, dfb_fast::dde_event_query};
  out->strict = DfbLaunchTable{dfb_strict::run_ode_fixed_shim, dfb_strict::run_ode_fixed_periodic_shim,
                               dfb_strict::run_ode_event_shim, nullptr, nullptr,
                               {dfb_strict::launch_general_g1}, 1, nullptr, nullptr};
  return 0;
}
