// Instantiations + launcher of the fixed-step ODE hot-path kernel.
// Compiled twice (see dev/common.cuh): STRICT (-fmad=false) and FAST.
#include "dev/ode_fixed.cuh"
#include "launch.hpp"
#include "host/launch_geometry.hpp"

#include <cstdlib>
#include <cstring>

namespace DFB_NS {

namespace {
constexpr unsigned long long kMaskRk4 = amask_bit(1, 0) | amask_bit(2, 1) | amask_bit(3, 2);
constexpr bool kPrescaled = (DFB_ARITH == 1);
constexpr int kBlock = 128;

// (members per thread, block size) from the ensemble size: see host/launch_geometry.hpp.  Occupancy of
// each variant is queried once per instantiation.
template <auto kernel1, auto kernel2>
dfb_host::LaunchGeometry pick_geometry(bool allow2, size_t n) {
  static int occ1[3] = {0, 0, 0}, occ2[3] = {0, 0, 0}, n_sm = 0;
  static const int blocks[3] = {128, 64, 32};
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    for (int b = 0; b < 3; ++b) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1[b], kernel1, blocks[b], 0);
      if (allow2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2[b], kernel2, blocks[b], 0);
    }
  }
  // one member per thread executes a few more non-FP64 instructions per member-step (tableau operands,
  // loop control and the shared time grid are not amortised over two members)
  dfb_host::GeometryCandidate cand[6];
  int nc = 0;
  const int force_tpt = dfb_host::env_int("DFB_ODE_TPT", 0);
  const int force_block = dfb_host::env_int("DFB_ODE_BLOCK", 0);
  for (int tpt = 2; tpt >= 1; --tpt) {
    if (tpt == 2 && !allow2) continue;
    if (force_tpt && force_tpt != tpt && !(tpt == 1 && !allow2)) continue;
    for (int b = 0; b < 3; ++b) {
      if (force_block && force_block != blocks[b]) continue;
      cand[nc++] = {blocks[b], tpt, tpt == 2 ? occ2[b] : occ1[b], tpt == 2 ? 1.0 : 1.012};
    }
  }
  if (nc == 0) cand[nc++] = {128, 1, occ1[0], 1.012};
  return dfb_host::choose_geometry(n, n_sm, cand, nc);
}

// TRACEABLE: also instantiate the variant that samples the trajectory (Trace output policy); kept to
// the two tableaux people integrate with (rktp64, rk4) to bound the number of instantiations — a
// traced run with another tableau takes the general kernel.
template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, bool TRACEABLE = false>
bool try_launch(int S_rt, unsigned long long amask, unsigned bmask, const OdeFixedArgs& A,
                cudaStream_t stream, int* rc) {
  if (S_rt != S) return false;
  if ((amask & ~AMASK) != 0ull || (bmask & ~BMASK) != 0u) return false;  // kernel must cover every non-zero
  const bool trace = A.trace_every > 0;
  if (trace && !TRACEABLE) return false;
  constexpr bool kTwo = Sys::N * (S + 3) <= 40;  // the doubled register footprint still fits
  constexpr int kTpt2 = kTwo ? 2 : 1;
  if constexpr (TRACEABLE) {
    if (trace) {
      const auto g = pick_geometry<ode_fixed_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, 1, true>,
                                   ode_fixed_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, kTpt2, true>>(kTwo, A.n_traj);
      if (g.tpt == 2)
        ode_fixed_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, kTpt2, true><<<unsigned(g.grid), g.block, 0, stream>>>(A);
      else
        ode_fixed_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, 1, true><<<unsigned(g.grid), g.block, 0, stream>>>(A);
      *rc = int(cudaGetLastError());
      return true;
    }
  }
  const auto g = pick_geometry<ode_fixed_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, 1>,
                               ode_fixed_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, kTpt2>>(kTwo, A.n_traj);
  if (g.tpt == 2)
    ode_fixed_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, kTpt2><<<unsigned(g.grid), g.block, 0, stream>>>(A);
  else
    ode_fixed_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, 1><<<unsigned(g.grid), g.block, 0, stream>>>(A);
  *rc = int(cudaGetLastError());
  return true;
}

// exact-sparsity variants first, then the full-lower-triangle catch-all
template <class Sys>
bool launch_common(int S, unsigned long long am, unsigned bm, const OdeFixedArgs& A,
                   cudaStream_t st, int* rc) {
  return try_launch<Sys, 7, amask_full_lower(7), 0x7Du, true>(S, am, bm, A, st, rc) ||  // rktp64
         try_launch<Sys, 7, amask_full_lower(7), 0x7Fu>(S, am, bm, A, st, rc) ||
         try_launch<Sys, 4, kMaskRk4, 0xFu, true>(S, am, bm, A, st, rc) ||        // rk4
         try_launch<Sys, 4, amask_full_lower(4), 0xFu>(S, am, bm, A, st, rc);
}
template <class Sys>
bool launch_all_shapes(int S, unsigned long long am, unsigned bm, const OdeFixedArgs& A,
                       cudaStream_t st, int* rc) {
  return launch_common<Sys>(S, am, bm, A, st, rc) ||
         try_launch<Sys, 1, 0ull, 0x1u>(S, am, bm, A, st, rc) ||
         try_launch<Sys, 2, amask_full_lower(2), 0x3u>(S, am, bm, A, st, rc) ||
         try_launch<Sys, 3, amask_full_lower(3), 0x7u>(S, am, bm, A, st, rc) ||
         try_launch<Sys, 5, amask_full_lower(5), 0x1Fu>(S, am, bm, A, st, rc);
}
}  // namespace

// ---- fixed step + Periodic callbacks (ode_fixed_periodic_kernel) -----------------------
namespace {
template <class Sys, int S, unsigned long long AMASK, unsigned BMASK>
bool try_launch_periodic(int S_rt, unsigned long long amask, unsigned bmask, const OdeFixedArgs& A,
                         cudaStream_t stream, int* rc) {
  if (S_rt != S) return false;
  if ((amask & ~AMASK) != 0ull || (bmask & ~BMASK) != 0u) return false;
  // Wide states (the 12-component Lyapunov system) take 218 registers uncapped: 8 resident warps
  // per SM, latency-bound (ncu: FP64 pipe 75 % active, 0.7 eligible warps per cycle).  Capped at 200
  // registers in 64-thread blocks (10 warps) the pipe saturates: 216.6 -> 208.9 ms on the 2^18-member
  // Lorenz spectrum batch; 168 registers (12 warps) gives the same time.  DFB_PERIODIC_MINB=1|3|5|6
  // selects the build (tuning).  The block itself (<= the launch bound) is sized from the ensemble so that a
  // shard of the batch (2^18 members over 8 GPUs = 1 024 warps for 148 SMs) spreads evenly over the SMs.
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  auto pick = [&](int max_block) {
    dfb_host::GeometryCandidate cand[3];
    int nc = 0;
    for (int b : {128, 64, 32})
      if (b <= max_block) cand[nc++] = {b, 1, 0, 1.0};
    return dfb_host::choose_geometry(A.n_traj, n_sm, cand, nc);
  };
  bool wide = false;
  if constexpr (Sys::N >= 8) {
    const char* e = getenv("DFB_PERIODIC_MINB");
    const int m = e ? atoi(e) : 5;
    wide = m == 3 || m == 4 || m == 5 || m == 6 || m == 7;
    const auto g128 = pick(128), g64 = pick(64);
    if (m == 4)
      ode_fixed_periodic_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, 1, 4><<<unsigned(g128.grid), g128.block, 0, stream>>>(A);
    else if (m == 7)
      ode_fixed_periodic_kernel<Sys, S, AMASK, BMASK, kPrescaled, 64, 1, 7><<<unsigned(g64.grid), g64.block, 0, stream>>>(A);
    else if (m == 3)
      ode_fixed_periodic_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, 1, 3><<<unsigned(g128.grid), g128.block, 0, stream>>>(A);
    else if (m == 5)
      ode_fixed_periodic_kernel<Sys, S, AMASK, BMASK, kPrescaled, 64, 1, 5><<<unsigned(g64.grid), g64.block, 0, stream>>>(A);
    else if (m == 6)
      ode_fixed_periodic_kernel<Sys, S, AMASK, BMASK, kPrescaled, 64, 1, 6><<<unsigned(g64.grid), g64.block, 0, stream>>>(A);
  }
  if (!wide) {
    const auto g = pick(128);
    ode_fixed_periodic_kernel<Sys, S, AMASK, BMASK, kPrescaled, kBlock, 1><<<unsigned(g.grid), g.block, 0, stream>>>(A);
  }
  *rc = int(cudaGetLastError());
  return true;
}
template <class Sys>
bool launch_periodic_common(int S, unsigned long long am, unsigned bm, const OdeFixedArgs& A,
                            cudaStream_t st, int* rc) {
  return try_launch_periodic<Sys, 7, amask_full_lower(7), 0x7Du>(S, am, bm, A, st, rc) ||  // rktp64
         try_launch_periodic<Sys, 4, amask_full_lower(4), 0xFu>(S, am, bm, A, st, rc);      // rk4 family
}
}  // namespace

int launch_ode_fixed_periodic(int system_id, int S, unsigned long long amask, unsigned bmask,
                              const OdeFixedArgs& A, cudaStream_t stream) {
  int rc = 0;
  bool ok = false;
#ifdef DFB_PLUGIN_SYSTEM
  (void)system_id;
  if constexpr (!DFB_PLUGIN_SYSTEM::uses_history && !DFB_PLUGIN_SYSTEM::reads_d)
    ok = launch_periodic_common<DFB_PLUGIN_SYSTEM>(S, amask, bmask, A, stream, &rc);
  return ok ? rc : -1;
#else
  switch (system_id) {
    case DFB_SYS_LORENZ_VARIATIONAL:
      ok = launch_periodic_common<SysLorenzVariational>(S, amask, bmask, A, stream, &rc);
      break;
    case DFB_SYS_LINEAR1: ok = launch_periodic_common<SysLinear1>(S, amask, bmask, A, stream, &rc); break;
    case DFB_SYS_LINEAR1_SIN: ok = launch_periodic_common<SysLinear1Sin>(S, amask, bmask, A, stream, &rc); break;
    case DFB_SYS_LINEAR3: ok = launch_periodic_common<SysLinear3>(S, amask, bmask, A, stream, &rc); break;
    case DFB_SYS_LINEAR3_MATRIX:
      ok = launch_periodic_common<SysLinear3Matrix>(S, amask, bmask, A, stream, &rc);
      break;
    case DFB_SYS_LORENZ: ok = launch_periodic_common<SysLorenz>(S, amask, bmask, A, stream, &rc); break;
    case DFB_SYS_EGG_BILLIARD:  // its callback calls stop_integration: exercises the per-lane stop
      ok = try_launch_periodic<SysEggBilliard, 7, amask_full_lower(7), 0x7Du>(S, amask, bmask, A, stream, &rc);
      break;
    default: break;
  }
  return ok ? rc : -1;
#endif
}

// returns 0 = launched, >0 = CUDA error code, -1 = no compiled kernel for this combination
int launch_ode_fixed(int system_id, int S, unsigned long long amask, unsigned bmask,
                     const OdeFixedArgs& A, cudaStream_t stream) {
  int rc = 0;
  bool ok = false;
#ifdef DFB_PLUGIN_SYSTEM
  (void)system_id;
  if constexpr (!DFB_PLUGIN_SYSTEM::uses_history && !DFB_PLUGIN_SYSTEM::reads_d)
    ok = launch_common<DFB_PLUGIN_SYSTEM>(S, amask, bmask, A, stream, &rc);
  return ok ? rc : -1;
#else
  switch (system_id) {
    case DFB_SYS_LORENZ: ok = launch_all_shapes<SysLorenz>(S, amask, bmask, A, stream, &rc); break;
    case DFB_SYS_LINEAR1: ok = launch_all_shapes<SysLinear1>(S, amask, bmask, A, stream, &rc); break;
    case DFB_SYS_HARMONIC: ok = launch_all_shapes<SysHarmonic>(S, amask, bmask, A, stream, &rc); break;
    case DFB_SYS_LORENZ_VARIATIONAL:
      ok = launch_common<SysLorenzVariational>(S, amask, bmask, A, stream, &rc);
      break;
    case DFB_SYS_LINEAR1_SIN: ok = launch_common<SysLinear1Sin>(S, amask, bmask, A, stream, &rc); break;
    case DFB_SYS_LINEAR3: ok = launch_common<SysLinear3>(S, amask, bmask, A, stream, &rc); break;
    case DFB_SYS_LINEAR3_MATRIX:
      ok = launch_common<SysLinear3Matrix>(S, amask, bmask, A, stream, &rc);
      break;
    default: break;
  }
  return ok ? rc : -1;
#endif
}

// Host shim used by api.cu: builds the argument block (pre-scaled rows are
// computed here with plain IEEE double multiplies) and launches.
namespace {
void fill_args(OdeFixedArgs& A, unsigned long long* amask_out, unsigned* bmask_out, const dfb_tableau& rk,
               double t0, double t1, double h, size_t n, const double* y0, const double* params,
               double* t_final, double* y_final, unsigned long long* steps,
               unsigned long long* rhs_evals, uint32_t* status) {
  std::memset(&A, 0, sizeof(A));
  unsigned long long amask = 0;
  unsigned bmask = 0;
  for (int i = 0; i < DFB_MAX_STAGES; ++i) {
    for (int j = 0; j < DFB_MAX_STAGES; ++j) {
      const double a = rk.a[i * DFB_MAX_STAGES + j];
      A.a[i * DFB_MAX_STAGES + j] = a;
      A.ha[i * DFB_MAX_STAGES + j] = h * a;
      if (i < rk.S && j < i && a != 0.0) amask |= amask_bit(i, j);
    }
    A.b[i] = rk.b[i];
    A.c[i] = rk.c[i];
    A.hb[i] = h * rk.b[i];
    A.hc[i] = rk.c[i] * h;
    if (i < rk.S && rk.b[i] != 0.0) bmask |= 1u << i;
  }
  A.t0 = t0;
  A.t1 = t1;
  A.h = h;
  A.n_traj = n;
  A.y0 = y0;
  A.params = params;
  A.t_final = t_final;
  A.y_final = y_final;
  A.steps = steps;
  A.rhs_evals = rhs_evals;
  A.status = status;
  *amask_out = amask;
  *bmask_out = bmask;
}
}  // namespace

int run_ode_fixed_shim(int system_id, const dfb_tableau& rk, double t0, double t1, double h,
                       size_t n, const double* y0, const double* params, double* t_final,
                       double* y_final, unsigned long long* steps, unsigned long long* rhs_evals,
                       uint32_t* status, int trace_every, int trace_capacity, double* trace_t,
                       double* trace_y, uint32_t* n_samples, cudaStream_t stream) {
  OdeFixedArgs A;
  unsigned long long amask = 0;
  unsigned bmask = 0;
  fill_args(A, &amask, &bmask, rk, t0, t1, h, n, y0, params, t_final, y_final, steps, rhs_evals, status);
  A.trace_every = (trace_t && trace_capacity > 0) ? trace_every : 0;
  A.trace_capacity = trace_capacity;
  A.trace_t = trace_t;
  A.trace_y = trace_y;
  A.n_samples = n_samples;
  return launch_ode_fixed(system_id, rk.S, amask, bmask, A, stream);
}

// Fixed step + Periodic locators (all `n_loc` entries of `loc` are DFB_LOC_PERIODIC, unfiltered).
int run_ode_fixed_periodic_shim(int system_id, const dfb_tableau& rk, double t0, double t1, double h,
                                size_t n, const double* y0, const double* params, double* t_final,
                                double* y_final, double* aux_final, unsigned long long* steps,
                                unsigned long long* events, unsigned long long* rhs_evals,
                                uint32_t* status, int n_loc, const dfb_locator* loc, int event_capacity,
                                double* ev_t, int32_t* ev_idx, uint32_t* n_events, cudaStream_t stream) {
  OdeFixedArgs A;
  unsigned long long amask = 0;
  unsigned bmask = 0;
  fill_args(A, &amask, &bmask, rk, t0, t1, h, n, y0, params, t_final, y_final, steps, rhs_evals, status);
  A.n_loc = n_loc;
  for (int l = 0; l < n_loc && l < DFB_MAX_LOCATORS; ++l) A.loc[l] = loc[l];
  A.event_capacity = ev_t ? event_capacity : 0;
  A.aux_final = aux_final;
  A.events = events;
  A.ev_t = ev_t;
  A.ev_idx = ev_idx;
  A.n_events = n_events;
  return launch_ode_fixed_periodic(system_id, rk.S, amask, bmask, A, stream);
}

}  // namespace DFB_NS
