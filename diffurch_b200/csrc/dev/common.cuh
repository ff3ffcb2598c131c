// Shared device/host declarations of the B200 ensemble integrator.
//
// Every translation unit under csrc/kernels_*.cu is compiled twice:
//   -DDFB_ARITH=0 -fmad=false  -> namespace dfb_strict (separately rounded mul/add,
//                                 the reference's operation order: bit-parity mode)
//   -DDFB_ARITH=1              -> namespace dfb_fast   (FMA contraction allowed)
#pragma once
#ifdef __CUDACC_RTC__
typedef struct CUstream_st* cudaStream_t;  // NVRTC: device code only; LaunchCfg below just needs the name
#else
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#endif

#include "../../../include/diffurch_b200.h"

#ifndef DFB_ARITH
#define DFB_ARITH 1
#endif
#if DFB_ARITH == 0
#define DFB_NS dfb_strict
#else
#define DFB_NS dfb_fast
#endif

#define DFB_DISCO_CAP 16  // per-trajectory discontinuity list slots (device)

// Kernel parameter block: lives in the constant bank (c[0x0]) as a
// __grid_constant__ argument, so tableau entries indexed with compile-time
// indices become direct constant operands of DFMA/DMUL.
struct DevProblem {
  dfb_tableau rk;
  dfb_locator loc[DFB_MAX_LOCATORS];
  dfb_auto_stepsize automatic;
  double t0, t1, h, max_delay;
  double disco_t[DFB_MAX_DISCO];
  int32_t disco_order[DFB_MAX_DISCO];
  int32_t n_disco;
  int32_t n_loc;
  int32_t history_id;
  int32_t stepsize_kind;
  int32_t trace_every, trace_capacity, event_capacity, ring_capacity;
  int32_t uses_disco;  // any PROPAGATION locator or initial_disco entry
  int32_t coeff_records;  // FAST: history records hold pre-combined interpolant coefficients (integrator.cuh)
  int32_t ode_history;    // an ODE problem whose locators may evaluate the state outside the step in flight
  int32_t on_start_cb;    // Solver::on_start_mut: callback id run on the initial State (0 = none)
  long long max_steps;
  size_t n_traj;
  // inputs (SoA, trajectory index fastest)
  const double* y0;      // [N][n]
  const double* params;  // [P][n]
  // outputs
  double* t_final;       // [n]
  double* y_final;       // [N][n]
  double* aux_final;     // [A][n]
  unsigned long long* steps;
  unsigned long long* rejected;
  unsigned long long* events;
  unsigned long long* rhs_evals;
  uint32_t* status;
  // trace [cap][n], [cap][N][n]
  double* trace_t;
  double* trace_y;
  uint32_t* n_samples;
  // event log [ecap][n]
  double* ev_t;
  int32_t* ev_idx;
  uint32_t* n_events;
  // delay-history ring (reference record format: t, y, k[S]) [cap][..][n]
  double* ring_t;
  double* ring_y;
  double* ring_k;
  // discontinuity list [DFB_DISCO_CAP][n]
  double* disco_buf_t;
  int32_t* disco_buf_ord;
};

// What a launcher needs besides the problem block.
struct LaunchCfg {
  cudaStream_t stream;
  int block;
  int grid;
  int locate_max_wait;  // step rounds a parked lane may wait for its warp's event batch
};
