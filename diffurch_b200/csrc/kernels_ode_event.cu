// Instantiations + launcher of the fixed-step ODE + located-events kernel (dev/ode_event.cuh).
// Compiled twice (see dev/common.cuh): STRICT (reference dense output, bit-identical event
// times) and FAST (pre-combined interpolant).
#include "dev/ode_event.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace DFB_NS {
namespace {
constexpr int kBlock = 128;
constexpr bool kFast = (DFB_ARITH == 1);

template <class Kernel>
void launch_grid(Kernel kernel, const OdeFixedArgs& A, int n_sm, cudaStream_t stream) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0);
  const size_t want = (A.n_traj + kBlock - 1) / kBlock;
  const size_t grid = std::min(want, size_t(per_sm > 0 ? per_sm : 1) * size_t(n_sm > 0 ? n_sm : 1));
  kernel<<<unsigned(grid ? grid : 1), kBlock, 0, stream>>>(A);
}

// Locator sets known at compile time (dev/ode_event.cuh, SPEC): the shipped event examples.  Any other set
// of plain sign-change locators runs the same kernel with the descriptors read at run time.
template <class Sys>
struct EventSpec {
  static constexpr int n_loc = 0;
  static constexpr unsigned spec = kEvSpecRuntime;
};
#ifndef DFB_PLUGIN_SYSTEM
template <>
struct EventSpec<SysBouncingBall> {  // examples/bouncing-ball.rs:27-32: on(zero(dx)), on_mut(below_zero(x), ..)
  static constexpr int n_loc = 2;
  static constexpr unsigned spec = ev_spec_byte(1, DFB_DET_ZERO, true) | (ev_spec_byte(2, DFB_DET_BELOW_ZERO, true) << 8);  // detector ids: dx = 1, x = 2
};
template <>
struct EventSpec<SysEggBilliard> {  // examples/egg-billiard.rs:22: on_mut(above_zero(boundary), ..)
  static constexpr int n_loc = 1;
  static constexpr unsigned spec = ev_spec_byte(1, DFB_DET_ABOVE_ZERO, true);
};
template <>
struct EventSpec<SysRelayMsign> {  // examples/ode-2-relay-msign.rs:21: on(zero(x))
  static constexpr int n_loc = 1;
  static constexpr unsigned spec = ev_spec_byte(1, DFB_DET_ZERO, true);
};
#endif
inline unsigned spec_of(const OdeFixedArgs& A) {
  unsigned v = 0u;
  for (int l = 0; l < A.n_loc && l < 4; ++l)
    v |= ev_spec_byte(A.loc[l].detector_id, A.loc[l].detection, A.loc[l].dedup != 0) << (8 * l);
  return v;
}

template <class Sys, int S, int I>
bool try_launch(int S_rt, int I_rt, const OdeFixedArgs& A, int n_sm, cudaStream_t stream, int* rc) {
  if (S_rt != S || I_rt != I) return false;
  constexpr unsigned long long AM = amask_full_lower(S);
  constexpr unsigned BM = (1u << S) - 1u;
  // every locator a plain Locator::{zero, above_zero, below_zero}: the SIMPLE profile (S = 7 only, to
  // bound the number of instantiations)
  bool simple = S == 7 && A.n_loc <= 2 && !getenv("DFB_EVENT_GENERIC_PROFILE");
  for (int l = 0; l < A.n_loc && simple; ++l) {
    const dfb_locator& L = A.loc[l];
    simple = L.kind == DFB_LOC_STATEFN && L.location == DFB_LOCATE_BISECTION && L.filter == DFB_FILTER_NONE &&
             (L.detection == DFB_DET_ZERO || L.detection == DFB_DET_ABOVE_ZERO || L.detection == DFB_DET_BELOW_ZERO);
  }
  if constexpr (S == 7) {
    const bool lean = A.trace_every == 0 && !getenv("DFB_EVENT_NO_LEAN");  // no Trace policy: LEAN build
    if constexpr (kFast && EventSpec<Sys>::n_loc > 0) {
      using ES = EventSpec<Sys>;
      bool match = simple && lean && A.n_loc == ES::n_loc && spec_of(A) == ES::spec && !getenv("DFB_EVENT_RUNTIME_SPEC");
      for (int l = 0; l < A.n_loc && match; ++l) match = A.loc[l].detector_id >= 0 && A.loc[l].detector_id < 16;
      if (match) {
        launch_grid(ode_event_kernel<Sys, S, I, AM, BM, kFast, ES::n_loc, kBlock, true, true, ES::spec>, A, n_sm, stream);
        *rc = int(cudaGetLastError());
        return true;
      }
    }
    if (simple && A.n_loc <= 1) {
      if (lean) launch_grid(ode_event_kernel<Sys, S, I, AM, BM, kFast, 1, kBlock, true, true>, A, n_sm, stream);
      else launch_grid(ode_event_kernel<Sys, S, I, AM, BM, kFast, 1, kBlock, true>, A, n_sm, stream);
      *rc = int(cudaGetLastError());
      return true;
    }
    if (simple && A.n_loc == 2) {
      if (lean) launch_grid(ode_event_kernel<Sys, S, I, AM, BM, kFast, 2, kBlock, true, true>, A, n_sm, stream);
      else launch_grid(ode_event_kernel<Sys, S, I, AM, BM, kFast, 2, kBlock, true>, A, n_sm, stream);
      *rc = int(cudaGetLastError());
      return true;
    }
  }
  if (A.n_loc <= 1) launch_grid(ode_event_kernel<Sys, S, I, AM, BM, kFast, 1, kBlock>, A, n_sm, stream);
  else if (A.n_loc == 2) launch_grid(ode_event_kernel<Sys, S, I, AM, BM, kFast, 2, kBlock>, A, n_sm, stream);
  else launch_grid(ode_event_kernel<Sys, S, I, AM, BM, kFast, DFB_MAX_LOCATORS, kBlock>, A, n_sm, stream);
  *rc = int(cudaGetLastError());
  return true;
}
}  // namespace

// returns 0 = launched, >0 = CUDA error code, -1 = no compiled kernel for this combination.
// `A` comes pre-filled by fill_event_args (kernels_ode_fixed.cu shares the argument block).
int launch_ode_event(int system_id, int S, int I, const OdeFixedArgs& A, int n_sm, cudaStream_t stream) {
  int rc = 0;
  bool ok = false;
#ifdef DFB_PLUGIN_SYSTEM
  (void)system_id;
  if constexpr (!DFB_PLUGIN_SYSTEM::uses_history && !DFB_PLUGIN_SYSTEM::reads_d)
    ok = try_launch<DFB_PLUGIN_SYSTEM, 7, 5>(S, I, A, n_sm, stream, &rc) ||
         try_launch<DFB_PLUGIN_SYSTEM, 5, 4>(S, I, A, n_sm, stream, &rc) ||
         try_launch<DFB_PLUGIN_SYSTEM, 4, 2>(S, I, A, n_sm, stream, &rc);
  return ok ? rc : -1;
#else
  switch (system_id) {
    case DFB_SYS_BOUNCING_BALL:
      ok = try_launch<SysBouncingBall, 7, 5>(S, I, A, n_sm, stream, &rc) ||
           try_launch<SysBouncingBall, 5, 4>(S, I, A, n_sm, stream, &rc) ||
           try_launch<SysBouncingBall, 4, 2>(S, I, A, n_sm, stream, &rc);
      break;
    case DFB_SYS_EGG_BILLIARD: ok = try_launch<SysEggBilliard, 7, 5>(S, I, A, n_sm, stream, &rc); break;
    case DFB_SYS_RELAY_MSIGN:
      ok = try_launch<SysRelayMsign, 7, 5>(S, I, A, n_sm, stream, &rc) ||
           try_launch<SysRelayMsign, 4, 2>(S, I, A, n_sm, stream, &rc);
      break;
    case DFB_SYS_LINEAR3: ok = try_launch<SysLinear3, 7, 5>(S, I, A, n_sm, stream, &rc); break;
    case DFB_SYS_HARMONIC: ok = try_launch<SysHarmonic, 7, 5>(S, I, A, n_sm, stream, &rc); break;
    default: break;
  }
  return ok ? rc : -1;
#endif
}

int run_ode_event_shim(int system_id, const dfb_tableau& rk, double t0, double t1, double h, long long max_steps,
                       size_t n, const double* y0, const double* params, double* t_final, double* y_final,
                       double* aux_final, unsigned long long* steps, unsigned long long* events,
                       unsigned long long* rhs_evals, unsigned long long* locate_iters, uint32_t* status,
                       int n_loc, const dfb_locator* loc,
                       int event_capacity, double* ev_t, int32_t* ev_idx, uint32_t* n_events, int trace_every,
                       int trace_capacity, double* trace_t, double* trace_y, uint32_t* n_samples,
                       unsigned long long* queue, int locate_batch, int locate_wait, int n_sm,
                       int on_start_cb, cudaStream_t stream) {
  OdeFixedArgs A;
  std::memset(&A, 0, sizeof(A));
  for (int i = 0; i < DFB_MAX_STAGES; ++i) {
    for (int j = 0; j < DFB_MAX_STAGES; ++j) A.a[i * DFB_MAX_STAGES + j] = rk.a[i * DFB_MAX_STAGES + j];
    A.b[i] = rk.b[i];
    A.c[i] = rk.c[i];
    for (int j = 0; j < DFB_MAX_INTERP; ++j) {
      A.bi[i * DFB_MAX_INTERP + j] = rk.bi[i * DFB_MAX_INTERP + j];
      // the FAST interpolant is p_prev + sum_{j>=1} c_j theta^j: needs b_i(0) = 0 (true of every rk.rs tableau)
      if (kFast && j == 0 && i < rk.S && rk.bi[i * DFB_MAX_INTERP] != 0.0) return -1;
    }
  }
  A.t0 = t0;
  A.t1 = t1;
  A.h = h;
  A.max_steps = max_steps;
  A.n_traj = n;
  A.y0 = y0;
  A.params = params;
  A.t_final = t_final;
  A.y_final = y_final;
  A.aux_final = aux_final;
  A.steps = steps;
  A.events = events;
  A.rhs_evals = rhs_evals;
  A.locate_iters = locate_iters;
  A.status = status;
  A.n_loc = n_loc;
  for (int l = 0; l < n_loc && l < DFB_MAX_LOCATORS; ++l) A.loc[l] = loc[l];
  A.event_capacity = ev_t ? event_capacity : 0;
  A.ev_t = ev_t;
  A.ev_idx = ev_idx;
  A.n_events = n_events;
  A.trace_every = trace_t ? trace_every : 0;
  A.trace_capacity = trace_capacity;
  A.trace_t = trace_t;
  A.trace_y = trace_y;
  A.n_samples = n_samples;
  A.queue = queue;
  A.locate_batch = locate_batch;
  A.locate_wait = locate_wait;
  A.on_start_cb = on_start_cb;
  A.inner_steps = 4;
  if (const char* e = std::getenv("DFB_EVENT_INNER")) A.inner_steps = std::atoi(e) > 0 ? std::atoi(e) : 1;
  return launch_ode_event(system_id, rk.S, rk.I, A, n_sm, stream);
}

}  // namespace DFB_NS
