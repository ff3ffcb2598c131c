// Interface between api.cu and the run-time compilation module (csrc/nvrtc_systems.cu).
#pragma once
#include "launch_table.hpp"

struct dfb_nvrtc_sysinfo {
  int n_state, n_params, n_aux, n_detectors, n_callbacks, uses_history, reads_d, pure_rhs, has_const_delay;
};

// Probe the functor `struct_name` defined by `source` (static facts read from NVRTC's PTX: no GPU needed)
// and fill the launch tables with shims that compile the needed kernel instantiation on first use.
int dfb_nvrtc_register(const char* source, const char* struct_name, int system_id, dfb_nvrtc_sysinfo* info,
                       DfbLaunchTable* fast, DfbLaunchTable* strict);
const char* dfb_nvrtc_last_error();
void dfb_nvrtc_get_stats(unsigned* compiled, unsigned* cache_hits, double* total_ms, double* last_ms);
int dfb_nvrtc_prebuild_ode_fixed(int system_id, int arith, const dfb_tableau* rk, int tpt, int trace);
