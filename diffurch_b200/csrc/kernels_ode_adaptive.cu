// Instantiations + launcher of the adaptive-step ODE hot-path kernel (FAST arithmetic
// only; STRICT goes through the general integrator, which keeps the reference's operation
// order and its libm `pow`).
#include "dev/ode_adaptive.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace DFB_NS {
namespace {
constexpr int kBlock = 128;

// trajectories per thread: measured on the 2^20-member Lorenz config, one slot per thread at
// 75 registers (24 resident warps/SM) beats two slots at 138 registers (209 ms vs 290 ms);
// DFB_ADAPT_TPT=2 selects the two-slot variant for tuning.
template <class Sys, int S>
constexpr int default_tpt() { return 1; }

// Persistent grid: every resident thread slot pulls trajectories from the work counter until it
// is empty, so the grid is sized by the device (resident blocks x SMs), not by the ensemble.
template <class Kernel>
void launch_grid(Kernel kernel, const OdeAdaptiveArgs& A, int tpt, int n_sm, cudaStream_t stream) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0);
  const size_t want = (A.n_traj + size_t(kBlock) * tpt - 1) / (size_t(kBlock) * tpt);
  const size_t grid = std::min(want, size_t(per_sm > 0 ? per_sm : 1) * size_t(n_sm > 0 ? n_sm : 1));
  kernel<<<unsigned(grid ? grid : 1), kBlock, 0, stream>>>(A);
}

template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, unsigned EMASK>
bool try_launch(int S_rt, unsigned long long am, unsigned bm, unsigned em, const OdeAdaptiveArgs& A,
                int n_sm, cudaStream_t stream, int* rc) {
  static_assert(Sys::pure_rhs, "the adaptive hot path needs a right-hand side that depends on (t, p, params) only");
  if (S_rt != S) return false;
  if ((am & ~AMASK) || (bm & ~BMASK) || (em & ~EMASK)) return false;
  if (A.trace_every > 0) {  // Trace policy: one-slot instantiations only
    if (A.p == 5) launch_grid(ode_adaptive_kernel<Sys, S, AMASK, BMASK, EMASK, kBlock, 1, 5, true>, A, 1, n_sm, stream);
    else launch_grid(ode_adaptive_kernel<Sys, S, AMASK, BMASK, EMASK, kBlock, 1, 0, true>, A, 1, n_sm, stream);
    *rc = int(cudaGetLastError());
    return true;
  }
  int tpt = default_tpt<Sys, S>();
  if (const char* e = getenv("DFB_ADAPT_TPT")) tpt = atoi(e) == 2 ? 2 : 1;
  if (Sys::N * (S + 3) > 40) tpt = 1;
  if (tpt == 2) {
    if constexpr (Sys::N * (S + 3) <= 40) {
      if (A.p == 5) launch_grid(ode_adaptive_kernel<Sys, S, AMASK, BMASK, EMASK, kBlock, 2, 5>, A, 2, n_sm, stream);
      else launch_grid(ode_adaptive_kernel<Sys, S, AMASK, BMASK, EMASK, kBlock, 2, 0>, A, 2, n_sm, stream);
    }
  } else {
    if (A.p == 5) launch_grid(ode_adaptive_kernel<Sys, S, AMASK, BMASK, EMASK, kBlock, 1, 5>, A, 1, n_sm, stream);
    else launch_grid(ode_adaptive_kernel<Sys, S, AMASK, BMASK, EMASK, kBlock, 1, 0>, A, 1, n_sm, stream);
  }
  *rc = int(cudaGetLastError());
  return true;
}

template <class Sys>
bool launch_common(int S, unsigned long long am, unsigned bm, unsigned em, const OdeAdaptiveArgs& A,
                   int n_sm, cudaStream_t st, int* rc) {
  return try_launch<Sys, 7, amask_full_lower(7), 0x7Du, 0x3Du>(S, am, bm, em, A, n_sm, st, rc) ||  // rktp64
         try_launch<Sys, 7, amask_full_lower(7), 0x7Fu, 0x7Fu>(S, am, bm, em, A, n_sm, st, rc) ||
         try_launch<Sys, 5, amask_full_lower(5), 0x1Fu, 0x1Fu>(S, am, bm, em, A, n_sm, st, rc) ||  // rk43
         try_launch<Sys, 4, amask_full_lower(4), 0xFu, 0xFu>(S, am, bm, em, A, n_sm, st, rc);      // rk4 family
}
template <class Sys>
bool launch_all_shapes(int S, unsigned long long am, unsigned bm, unsigned em, const OdeAdaptiveArgs& A,
                       int n_sm, cudaStream_t st, int* rc) {
  return launch_common<Sys>(S, am, bm, em, A, n_sm, st, rc) ||
         try_launch<Sys, 3, amask_full_lower(3), 0x7u, 0x7u>(S, am, bm, em, A, n_sm, st, rc) ||
         try_launch<Sys, 2, amask_full_lower(2), 0x3u, 0x3u>(S, am, bm, em, A, n_sm, st, rc) ||
         try_launch<Sys, 1, 0ull, 0x1u, 0x1u>(S, am, bm, em, A, n_sm, st, rc);
}
}  // namespace

// returns 0 = launched, >0 = CUDA error code, -1 = no compiled kernel for this combination
int run_ode_adaptive_shim(int system_id, const dfb_tableau& rk, const dfb_auto_stepsize& ctl, double t0,
                          double t1, size_t n, const double* y0, const double* params, double* t_final,
                          double* y_final, unsigned long long* steps, unsigned long long* rejected,
                          unsigned long long* rhs_evals, uint32_t* status, unsigned long long* queue,
                          int n_sm, int trace_every, int trace_capacity, double* trace_t, double* trace_y,
                          uint32_t* n_samples, cudaStream_t stream) {
  OdeAdaptiveArgs A;
  std::memset(&A, 0, sizeof(A));
  unsigned long long am = 0;
  unsigned bm = 0, em = 0;
  for (int i = 0; i < DFB_MAX_STAGES; ++i) {
    for (int j = 0; j < DFB_MAX_STAGES; ++j) {
      const double a = rk.a[i * DFB_MAX_STAGES + j];
      A.a[i * DFB_MAX_STAGES + j] = a;
      if (i < rk.S && j < i && a != 0.0) am |= amask_bit(i, j);
    }
    A.b[i] = rk.b[i];
    A.c[i] = rk.c[i];
    A.e[i] = rk.b2[i] - rk.b[i];
    if (i < rk.S && rk.b[i] != 0.0) bm |= 1u << i;
    if (i < rk.S && A.e[i] != 0.0) em |= 1u << i;
  }
  A.ctl = ctl;
  A.t0 = t0;
  A.t1 = t1;
  A.p = int(ctl.order + 1u);
  A.inv_p = 1.0 / double(ctl.order + 1u);
  A.inv_pf = float(A.inv_p);
  A.n_traj = n;
  A.y0 = y0;
  A.params = params;
  A.t_final = t_final;
  A.y_final = y_final;
  A.steps = steps;
  A.rejected = rejected;
  A.rhs_evals = rhs_evals;
  A.status = status;
  A.queue = queue;
  A.trace_every = (trace_t && trace_capacity > 0) ? trace_every : 0;
  A.trace_capacity = trace_capacity;
  A.trace_t = trace_t;
  A.trace_y = trace_y;
  A.n_samples = n_samples;
  int rc = 0;
  bool ok = false;
#ifdef DFB_PLUGIN_SYSTEM
  (void)system_id;
  if constexpr (!DFB_PLUGIN_SYSTEM::uses_history && DFB_PLUGIN_SYSTEM::pure_rhs)
    ok = launch_common<DFB_PLUGIN_SYSTEM>(rk.S, am, bm, em, A, n_sm, stream, &rc);
  return ok ? rc : -1;
#else
  switch (system_id) {
    case DFB_SYS_LORENZ: ok = launch_all_shapes<SysLorenz>(rk.S, am, bm, em, A, n_sm, stream, &rc); break;
    case DFB_SYS_HARMONIC: ok = launch_all_shapes<SysHarmonic>(rk.S, am, bm, em, A, n_sm, stream, &rc); break;
    case DFB_SYS_LINEAR1: ok = launch_all_shapes<SysLinear1>(rk.S, am, bm, em, A, n_sm, stream, &rc); break;
    case DFB_SYS_LINEAR1_SIN: ok = launch_common<SysLinear1Sin>(rk.S, am, bm, em, A, n_sm, stream, &rc); break;
    case DFB_SYS_LINEAR3: ok = launch_common<SysLinear3>(rk.S, am, bm, em, A, n_sm, stream, &rc); break;
    case DFB_SYS_POWER_DECAY: ok = launch_common<SysPowerDecay>(rk.S, am, bm, em, A, n_sm, stream, &rc); break;
    case DFB_SYS_LORENZ_VARIATIONAL:
      ok = launch_common<SysLorenzVariational>(rk.S, am, bm, em, A, n_sm, stream, &rc);
      break;
    default: break;
  }
  return ok ? rc : -1;
#endif
}

}  // namespace DFB_NS
