// Host-side launcher signatures of every kernel family, one set per arithmetic namespace
// (dfb_strict / dfb_fast), and the table api.cu dispatches through.  The built-in systems fill
// the table with the launchers linked into libdiffurch_b200.so; a user-system plugin
// (csrc/plugin_entry.cu, built by diffurch_b200.user.build_user_system) hands over a table of
// its own launchers, instantiated from the same kernel templates for the user's device functor.
#pragma once
#include "dev/common.cuh"

typedef int (*DfbOdeFixedFn)(int system_id, const dfb_tableau& rk, double t0, double t1, double h, size_t n,
                             const double* y0, const double* params, double* t_final, double* y_final,
                             unsigned long long* steps, unsigned long long* rhs_evals, uint32_t* status,
                             int trace_every, int trace_capacity, double* trace_t, double* trace_y,
                             uint32_t* n_samples, cudaStream_t stream);
typedef int (*DfbOdePeriodicFn)(int system_id, const dfb_tableau& rk, double t0, double t1, double h, size_t n,
                                const double* y0, const double* params, double* t_final, double* y_final,
                                double* aux_final, unsigned long long* steps, unsigned long long* events,
                                unsigned long long* rhs_evals, uint32_t* status, int n_loc,
                                const dfb_locator* loc, int event_capacity, double* ev_t, int32_t* ev_idx,
                                uint32_t* n_events, cudaStream_t stream);
typedef int (*DfbOdeEventFn)(int system_id, const dfb_tableau& rk, double t0, double t1, double h,
                             long long max_steps, size_t n, const double* y0, const double* params,
                             double* t_final, double* y_final, double* aux_final, unsigned long long* steps,
                             unsigned long long* events, unsigned long long* rhs_evals,
                             unsigned long long* locate_iters, uint32_t* status,
                             int n_loc, const dfb_locator* loc, int event_capacity, double* ev_t,
                             int32_t* ev_idx, uint32_t* n_events, int trace_every, int trace_capacity,
                             double* trace_t, double* trace_y, uint32_t* n_samples, unsigned long long* queue,
                             int locate_batch, int locate_wait, int n_sm, int on_start_cb, cudaStream_t stream);
typedef int (*DfbOdeAdaptiveFn)(int system_id, const dfb_tableau& rk, const dfb_auto_stepsize& ctl, double t0,
                                double t1, size_t n, const double* y0, const double* params, double* t_final,
                                double* y_final, unsigned long long* steps, unsigned long long* rejected,
                                unsigned long long* rhs_evals, uint32_t* status, unsigned long long* queue,
                                int n_sm, int trace_every, int trace_capacity, double* trace_t, double* trace_y,
                                uint32_t* n_samples, cudaStream_t stream);
typedef int (*DfbDdeFixedFn)(int system_id, const dfb_tableau& rk, double t0, double t1, double h,
                             double max_delay, int history_id, int ring_capacity, size_t n, const double* y0,
                             const double* params, double* ring, double* t_final, double* y_final,
                             unsigned long long* steps, unsigned long long* rhs_evals, uint32_t* status,
                             int trace_every, int trace_capacity, double* trace_t, double* trace_y,
                             uint32_t* n_samples, cudaStream_t stream);
typedef int (*DfbDdeEventFn)(int system_id, const dfb_tableau& rk, double t0, double t1, double h, double max_delay,
                             long long max_steps, int history_id, int ring_capacity, unsigned ring_warps,
                             double* ring, size_t n, const double* y0, const double* params, double* t_final,
                             double* y_final, double* aux_final, unsigned long long* steps,
                             unsigned long long* events, unsigned long long* rhs_evals,
                             unsigned long long* locate_iters, uint32_t* status, int n_loc, const dfb_locator* loc,
                             int n_disco, const double* disco_t, const int32_t* disco_order, int event_capacity,
                             double* ev_t, int32_t* ev_idx, uint32_t* n_events, int trace_every,
                             int trace_capacity, double* trace_t, double* trace_y, uint32_t* n_samples,
                             unsigned long long* queue, int locate_batch, int locate_wait, int n_sm,
                             int on_start_cb, cudaStream_t stream);
typedef int (*DfbGeneralFn)(const DevProblem& P, int system_id, bool has_loc, const LaunchCfg& c,
                            int locate_batch_threshold);

#define DFB_MAX_GENERAL_GROUPS 5

// One table per arithmetic contract.  Null entries = that family is not compiled for the
// contract (the adaptive and DDE hot paths are FAST only).
struct DfbLaunchTable {
  DfbOdeFixedFn ode_fixed;
  DfbOdePeriodicFn ode_periodic;
  DfbOdeEventFn ode_event;
  DfbOdeAdaptiveFn ode_adaptive;
  DfbDdeFixedFn dde_fixed;
  DfbGeneralFn general[DFB_MAX_GENERAL_GROUPS];
  int n_general;
  DfbDdeEventFn dde_event;  // FAST only
  int (*dde_event_query)(int system_id, int S, int I);
};

#define DFB_DECLARE_SHIMS(NS)                                                                        \
  namespace NS {                                                                                     \
  int run_ode_fixed_shim(int, const dfb_tableau&, double, double, double, size_t, const double*,    \
                         const double*, double*, double*, unsigned long long*, unsigned long long*, \
                         uint32_t*, int, int, double*, double*, uint32_t*, cudaStream_t);            \
  int run_ode_fixed_periodic_shim(int, const dfb_tableau&, double, double, double, size_t,          \
                                  const double*, const double*, double*, double*, double*,          \
                                  unsigned long long*, unsigned long long*, unsigned long long*,    \
                                  uint32_t*, int, const dfb_locator*, int, double*, int32_t*,       \
                                  uint32_t*, cudaStream_t);                                          \
  int run_ode_event_shim(int, const dfb_tableau&, double, double, double, long long, size_t,        \
                         const double*, const double*, double*, double*, double*,                   \
                         unsigned long long*, unsigned long long*, unsigned long long*,             \
                         unsigned long long*, uint32_t*,                                            \
                         int, const dfb_locator*, int, double*, int32_t*, uint32_t*, int, int,      \
                         double*, double*, uint32_t*, unsigned long long*, int, int, int, int,      \
                         cudaStream_t);                                                              \
  }

// What a plugin's `dfb_plugin_describe` fills in (csrc/plugin_entry.cu).
#define DFB_PLUGIN_ABI 7
struct dfb_plugin_desc {
  int plugin_abi;
  int n_state, n_params, n_aux, n_detectors, n_callbacks, uses_history;
  int dde_hot_ok;  // constant-delay system the coefficient-ring kernel can take
  // layout check of every struct that crosses the plugin boundary by reference
  unsigned sizeof_devproblem, sizeof_locator, sizeof_tableau, sizeof_auto_stepsize;
  const char* name;
  DfbLaunchTable fast, strict;
};
typedef int (*DfbPluginDescribeFn)(dfb_plugin_desc*);
