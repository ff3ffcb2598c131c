// Instantiations + launchers of the general integrator, split in groups
// (-DDFB_GROUP=1..5) so the build can compile them in parallel, and compiled
// twice per group for the two arithmetic contracts (see dev/common.cuh).
#include "dev/integrator.cuh"
#include "launch.hpp"

#ifndef DFB_GROUP
#error "define DFB_GROUP"
#endif

namespace DFB_NS {
namespace {
template <class Sys, int S, int I, bool RT = false>
int launch_one(const DevProblem& P, bool has_loc, const LaunchCfg& c, int thr) {
  const int wait = c.locate_max_wait;
  if (has_loc)
    integrate_general_kernel<Sys, S, I, true, RT><<<c.grid, c.block, 0, c.stream>>>(P, thr, wait);
  else
    integrate_general_kernel<Sys, S, I, false, RT><<<c.grid, c.block, 0, c.stream>>>(P, thr, wait);
  return int(cudaGetLastError());
}
}  // namespace

#define SHAPE(Sys, S_, I_) \
  if (S == S_ && I == I_) return launch_one<Sys, S_, I_>(P, has_loc, c, thr);
// Catch-all: the maximal-shape instantiation with run-time stage / interpolant counts
// (integrator.cuh, ParamTab<.., RT = true>) takes any tableau the exact shapes above it did not.
#define CATCHALL(Sys) return launch_one<Sys, DFB_MAX_STAGES, DFB_MAX_INTERP, true>(P, has_loc, c, thr);
// ODE problem that needs the reference's retained steps (RegulaFalsi, see Integrator<.., ODEHIST>)
#define ODEHIST(Sys)                                                                                  \
  if (P.ode_history) {                                                                                \
    if constexpr (!Sys::uses_history) {                                                               \
      integrate_general_kernel<Sys, DFB_MAX_STAGES, DFB_MAX_INTERP, true, true, true>                 \
          <<<c.grid, c.block, 0, c.stream>>>(P, thr, c.locate_max_wait);                              \
      return int(cudaGetLastError());                                                                 \
    }                                                                                                 \
  }
#define ALL_SHAPES(Sys) SHAPE(Sys, 7, 5) SHAPE(Sys, 5, 4) SHAPE(Sys, 4, 2) CATCHALL(Sys)

#define DFB_CAT2(a, b) a##b
#define DFB_CAT(a, b) DFB_CAT2(a, b)

// returns 0 = launched, >0 = CUDA error, -1 = (system, shape) not compiled in this group
int DFB_CAT(launch_general_g, DFB_GROUP)(const DevProblem& P, int system_id, bool has_loc,
                                         const LaunchCfg& c, int thr) {
  const int S = P.rk.S, I = P.rk.I;
#ifdef DFB_PLUGIN_SYSTEM
  (void)system_id;
  ODEHIST(DFB_PLUGIN_SYSTEM)
  SHAPE(DFB_PLUGIN_SYSTEM, 7, 5) SHAPE(DFB_PLUGIN_SYSTEM, 4, 2) CATCHALL(DFB_PLUGIN_SYSTEM)
#else
  switch (system_id) {
#if DFB_GROUP == 1
    case DFB_SYS_LORENZ: ODEHIST(SysLorenz) ALL_SHAPES(SysLorenz)
    case DFB_SYS_LORENZ_VARIATIONAL: ODEHIST(SysLorenzVariational)
      SHAPE(SysLorenzVariational, 7, 5) SHAPE(SysLorenzVariational, 4, 2) CATCHALL(SysLorenzVariational)
#elif DFB_GROUP == 2
    case DFB_SYS_LINEAR1: ODEHIST(SysLinear1) ALL_SHAPES(SysLinear1)
    case DFB_SYS_LINEAR1_SIN: ODEHIST(SysLinear1Sin) SHAPE(SysLinear1Sin, 7, 5) CATCHALL(SysLinear1Sin)
    case DFB_SYS_LINEAR3: ODEHIST(SysLinear3) SHAPE(SysLinear3, 7, 5) SHAPE(SysLinear3, 4, 2) CATCHALL(SysLinear3)
    case DFB_SYS_LINEAR3_MATRIX: ODEHIST(SysLinear3Matrix) SHAPE(SysLinear3Matrix, 7, 5) CATCHALL(SysLinear3Matrix)
#elif DFB_GROUP == 3
    case DFB_SYS_HARMONIC: ODEHIST(SysHarmonic) ALL_SHAPES(SysHarmonic)
    case DFB_SYS_POWER_DECAY: ODEHIST(SysPowerDecay) SHAPE(SysPowerDecay, 7, 5) CATCHALL(SysPowerDecay)
    case DFB_SYS_CONSTANT3: ODEHIST(SysConstant3) SHAPE(SysConstant3, 7, 5) SHAPE(SysConstant3, 1, 2) CATCHALL(SysConstant3)
    case DFB_SYS_TIME3: ODEHIST(SysTime3) SHAPE(SysTime3, 7, 5) SHAPE(SysTime3, 2, 2) CATCHALL(SysTime3)
    case DFB_SYS_ZERO1: ODEHIST(SysZero1) SHAPE(SysZero1, 7, 5) SHAPE(SysZero1, 4, 2) SHAPE(SysZero1, 1, 2) CATCHALL(SysZero1)
#elif DFB_GROUP == 4
    case DFB_SYS_DDE_LINEAR: ODEHIST(SysDdeLinear) SHAPE(SysDdeLinear, 7, 5) SHAPE(SysDdeLinear, 5, 4) SHAPE(SysDdeLinear, 4, 2) CATCHALL(SysDdeLinear)
    case DFB_SYS_NDDE_LINEAR: ODEHIST(SysNddeLinear) SHAPE(SysNddeLinear, 7, 5) SHAPE(SysNddeLinear, 5, 4) CATCHALL(SysNddeLinear)
    case DFB_SYS_DDE_RELAY_CLAMP: ODEHIST(SysDdeRelayClamp) SHAPE(SysDdeRelayClamp, 7, 5) CATCHALL(SysDdeRelayClamp)
#elif DFB_GROUP == 5
    case DFB_SYS_BOUNCING_BALL: ODEHIST(SysBouncingBall) SHAPE(SysBouncingBall, 7, 5) SHAPE(SysBouncingBall, 5, 4) SHAPE(SysBouncingBall, 4, 2) CATCHALL(SysBouncingBall)
    case DFB_SYS_EGG_BILLIARD: ODEHIST(SysEggBilliard) SHAPE(SysEggBilliard, 7, 5) CATCHALL(SysEggBilliard)
    case DFB_SYS_RELAY_MSIGN: ODEHIST(SysRelayMsign) SHAPE(SysRelayMsign, 7, 5) CATCHALL(SysRelayMsign)
#endif
    default: break;
  }
  (void)S;
  (void)I;
  return -1;
#endif
}

}  // namespace DFB_NS
