// Instantiations + launcher of the DDE/NDDE + located-events + propagation kernel
// (dev/dde_event.cuh).  FAST arithmetic only: STRICT keeps the reference's record format and
// operation order in the general integrator.
#include "dev/dde_event.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace DFB_NS {
namespace {
constexpr int kBlock = 128;

#ifndef DFB_DDE_EV_MINB
// 4 blocks of 128 threads per SM (128 registers, a few spilled doubles) beat 3 blocks at 154 registers by
// 31 % on the relay-clamp ensemble: the kernel is latency-bound, throughput follows the resident warps
#define DFB_DDE_EV_MINB(N) ((N) <= 2 ? 4 : 2)
#endif
template <class Sys>
constexpr int min_blocks() { return DFB_DDE_EV_MINB(Sys::N); }

template <class Kernel>
int launch_grid(Kernel kernel, DdeEventArgs& A, int n_sm, cudaStream_t stream, int block = kBlock) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0);
  const size_t want = (A.n_traj + block - 1) / block;
  size_t grid = std::min(want, size_t(per_sm > 0 ? per_sm : 1) * size_t(n_sm > 0 ? n_sm : 1));
  grid = std::min(grid, size_t(A.ring_warps) * 32 / block);  // one ring column per grid thread
  if (grid == 0) return -1;
  kernel<<<unsigned(grid), block, 0, stream>>>(A);
  return int(cudaGetLastError());
}

// FULL = instantiate every compile-time locator-kind combination of the SIMPLE profile (the shipped
// event example); otherwise only the combinations the reference's tests use, the rest read the kinds
// from the descriptor (KINDS = -1).
template <class Sys, int S, int I, bool FULL>
bool try_launch(int S_rt, int I_rt, DdeEventArgs& A, int n_sm, cudaStream_t stream, int* rc) {
  if (S_rt != S || I_rt != I) return false;
  bool has_prop = A.n_disco > 0;
  unsigned kinds = 0u;
  for (int l = 0; l < A.n_loc; ++l)
    if (A.loc[l].kind == DFB_LOC_PROPAGATION) {
      has_prop = true;
      kinds |= 1u << l;
    }
  // SIMPLE profile (see dev/dde_event.cuh): unfiltered propagation / sign-change + Bisection locators only
  bool simple = A.n_loc <= 2 && !getenv("DFB_EVENT_GENERIC_PROFILE");
  for (int l = 0; l < A.n_loc && simple; ++l) {
    const dfb_locator& L = A.loc[l];
    const bool sign = L.kind == DFB_LOC_STATEFN && L.location == DFB_LOCATE_BISECTION &&
                      (L.detection == DFB_DET_ZERO || L.detection == DFB_DET_ABOVE_ZERO || L.detection == DFB_DET_BELOW_ZERO);
    simple = (sign || L.kind == DFB_LOC_PROPAGATION) && L.filter == DFB_FILTER_NONE;
  }
  constexpr int MB = min_blocks<Sys>();
  const bool known = !getenv("DFB_EVENT_RUNTIME_KINDS");
  if (!simple) {
    *rc = launch_grid(dde_event_kernel<Sys, S, I, DFB_MAX_LOCATORS, true, false, -1, kBlock, MB>, A, n_sm, stream);
  } else if (A.n_loc <= 1 && !has_prop) {
    *rc = launch_grid(dde_event_kernel<Sys, S, I, 1, false, true, 0, kBlock, MB>, A, n_sm, stream);
  } else if (A.n_loc <= 1) {
    *rc = launch_grid(dde_event_kernel<Sys, S, I, 1, true, true, 1, kBlock, MB>, A, n_sm, stream);
  } else if (known && kinds == 1u) {  // with_const_delay(..) then on(Locator::zero(..)): dde-relay-clamp.rs:30-35
    if constexpr (FULL) {  // DFB_DDE_EV_MINB=3: the 154-register build without spills, 12 warps per SM (tuning)
      if (const char* e = getenv("DFB_DDE_EV_MINB")) {
        if (atoi(e) == 3) {
          *rc = launch_grid(dde_event_kernel<Sys, S, I, 2, true, true, 1, kBlock, 3>, A, n_sm, stream);
          return true;
        }

      }
    }
    *rc = launch_grid(dde_event_kernel<Sys, S, I, 2, true, true, 1, kBlock, MB>, A, n_sm, stream);
  } else if (known && kinds == 3u) {  // two propagated delays: delay_propagation.rs:28-51
    *rc = launch_grid(dde_event_kernel<Sys, S, I, 2, true, true, 3, kBlock, MB>, A, n_sm, stream);
  } else {
    if constexpr (FULL) {
      if (known && kinds == 0u) {
        *rc = launch_grid(dde_event_kernel<Sys, S, I, 2, false, true, 0, kBlock, MB>, A, n_sm, stream);
        return true;
      }
      if (known && kinds == 2u) {
        *rc = launch_grid(dde_event_kernel<Sys, S, I, 2, true, true, 2, kBlock, MB>, A, n_sm, stream);
        return true;
      }
    }
    *rc = launch_grid(dde_event_kernel<Sys, S, I, 2, true, true, -1, kBlock, MB>, A, n_sm, stream);
  }
  return true;
}
// compiled (system, tableau shape) pairs: rktp64 (the default, solver.rs:88) for the delay systems, the rk4
// family for the relay-clamp example and euler / rk4 for the propagation tests (tests/delay_propagation.rs)
constexpr bool shape_compiled(int system_id, int S, int I) {
  if (S == 7 && I == 5) return system_id != DFB_SYS_ZERO1;
  if (S == 4 && I == 2) return system_id == DFB_SYS_DDE_RELAY_CLAMP || system_id == DFB_SYS_ZERO1 || system_id >= DFB_SYS_USER_BASE;
  if (S == 1 && I == 2) return system_id == DFB_SYS_ZERO1;
  return false;
}
}  // namespace

// 1 = a dde_event_kernel instantiation exists for (system, tableau shape): api.cu asks before it
// sizes the coefficient ring (dfb_create), so an unsupported combination gets the general rings
int dde_event_query(int system_id, int S, int I) {
  if (!shape_compiled(system_id, S, I)) return 0;
#ifdef DFB_PLUGIN_SYSTEM
  (void)system_id;
  return (DFB_PLUGIN_SYSTEM::uses_history && !DFB_PLUGIN_SYSTEM::reads_d) ? 1 : 0;
#else
  return (system_id == DFB_SYS_DDE_RELAY_CLAMP || system_id == DFB_SYS_DDE_LINEAR ||
          system_id == DFB_SYS_NDDE_LINEAR || system_id == DFB_SYS_ZERO1) ? 1 : 0;
#endif
}

// doubles of ring per (lane, slot)
size_t dde_event_record_doubles(int n_state, int I) { return size_t(1 + n_state * I); }

// returns 0 = launched, >0 = CUDA error code, -1 = no compiled kernel for this combination
int run_dde_event_shim(int system_id, const dfb_tableau& rk, double t0, double t1, double h, double max_delay,
                       long long max_steps, int history_id, int ring_capacity, unsigned ring_warps, double* ring,
                       size_t n, const double* y0, const double* params, double* t_final, double* y_final,
                       double* aux_final, unsigned long long* steps, unsigned long long* events,
                       unsigned long long* rhs_evals, unsigned long long* locate_iters, uint32_t* status,
                       int n_loc, const dfb_locator* loc, int n_disco, const double* disco_t,
                       const int32_t* disco_order, int event_capacity, double* ev_t, int32_t* ev_idx,
                       uint32_t* n_events, int trace_every, int trace_capacity, double* trace_t, double* trace_y,
                       uint32_t* n_samples, unsigned long long* queue, int locate_batch, int locate_wait, int n_sm,
                       int on_start_cb, cudaStream_t stream) {
  DdeEventArgs A;
  std::memset(&A, 0, sizeof(A));
  for (int i = 0; i < DFB_MAX_STAGES; ++i) {
    for (int j = 0; j < DFB_MAX_STAGES; ++j) A.a[i * DFB_MAX_STAGES + j] = rk.a[i * DFB_MAX_STAGES + j];
    A.b[i] = rk.b[i];
    A.c[i] = rk.c[i];
    for (int j = 0; j < DFB_MAX_INTERP; ++j) {
      A.bi[i * DFB_MAX_INTERP + j] = rk.bi[i * DFB_MAX_INTERP + j];
      // the pre-combined interpolant is y_r + sum_{j>=1} d_j s^j: needs b_i(0) = 0 (true of every rk.rs tableau)
      if (j == 0 && i < rk.S && rk.bi[i * DFB_MAX_INTERP] != 0.0) return -1;
    }
  }
  A.t0 = t0;
  A.t1 = t1;
  A.h = h;
  A.max_delay = max_delay;
  A.max_steps = max_steps;
  A.history_id = history_id;
  A.ring_capacity = ring_capacity;
  A.rk_order = int(rk.order);
  A.n_disco = n_disco;
  for (int q = 0; q < n_disco && q < DFB_MAX_DISCO; ++q) {
    A.disco_t[q] = disco_t[q];
    A.disco_order[q] = disco_order[q];
  }
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
  A.ring = ring;
  A.ring_warps = ring_warps;
  A.prefetch_mode = 1;
  if (const char* e = std::getenv("DFB_DDE_EV_PREFETCH_MODE")) A.prefetch_mode = std::atoi(e);
  int rc = 0;
  bool ok = false;
#ifdef DFB_PLUGIN_SYSTEM
  (void)system_id;
  if constexpr (DFB_PLUGIN_SYSTEM::uses_history && !DFB_PLUGIN_SYSTEM::reads_d)
    ok = try_launch<DFB_PLUGIN_SYSTEM, 7, 5, false>(rk.S, rk.I, A, n_sm, stream, &rc) ||
         try_launch<DFB_PLUGIN_SYSTEM, 4, 2, false>(rk.S, rk.I, A, n_sm, stream, &rc);
  return ok ? rc : -1;
#else
  switch (system_id) {
    case DFB_SYS_DDE_RELAY_CLAMP:
      ok = try_launch<SysDdeRelayClamp, 7, 5, true>(rk.S, rk.I, A, n_sm, stream, &rc) ||
           try_launch<SysDdeRelayClamp, 4, 2, false>(rk.S, rk.I, A, n_sm, stream, &rc);
      break;
    case DFB_SYS_DDE_LINEAR: ok = try_launch<SysDdeLinear, 7, 5, false>(rk.S, rk.I, A, n_sm, stream, &rc); break;
    case DFB_SYS_NDDE_LINEAR: ok = try_launch<SysNddeLinear, 7, 5, false>(rk.S, rk.I, A, n_sm, stream, &rc); break;
    case DFB_SYS_ZERO1:  // an ODE with propagated discontinuities (tests/delay_propagation.rs)
      ok = try_launch<SysZero1, 4, 2, false>(rk.S, rk.I, A, n_sm, stream, &rc) ||
           try_launch<SysZero1, 1, 2, false>(rk.S, rk.I, A, n_sm, stream, &rc);
      break;
    default: break;
  }
  return ok ? rc : -1;
#endif
}

}  // namespace DFB_NS
