// Launcher entry points exported by the kernel translation units, one set per
// arithmetic namespace (dfb_strict / dfb_fast).
#pragma once
#include "dev/common.cuh"

#define DFB_DECLARE_LAUNCHERS(NS)                                                               \
  namespace NS {                                                                                \
  int launch_general_g1(const DevProblem& P, int system_id, bool has_loc, const LaunchCfg& c,   \
                        int locate_batch_threshold);                                            \
  int launch_general_g2(const DevProblem& P, int system_id, bool has_loc, const LaunchCfg& c,   \
                        int locate_batch_threshold);                                            \
  int launch_general_g3(const DevProblem& P, int system_id, bool has_loc, const LaunchCfg& c,   \
                        int locate_batch_threshold);                                            \
  int launch_general_g4(const DevProblem& P, int system_id, bool has_loc, const LaunchCfg& c,   \
                        int locate_batch_threshold);                                            \
  int launch_general_g5(const DevProblem& P, int system_id, bool has_loc, const LaunchCfg& c,   \
                        int locate_batch_threshold);                                            \
  }

DFB_DECLARE_LAUNCHERS(dfb_strict)
DFB_DECLARE_LAUNCHERS(dfb_fast)
