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
template <class Sys, int S, int I>
int launch_one(const DevProblem& P, bool has_loc, const LaunchCfg& c, int thr) {
  const int wait = c.locate_max_wait;
  if (has_loc)
    integrate_general_kernel<Sys, S, I, true><<<c.grid, c.block, 0, c.stream>>>(P, thr, wait);
  else
    integrate_general_kernel<Sys, S, I, false><<<c.grid, c.block, 0, c.stream>>>(P, thr, wait);
  return int(cudaGetLastError());
}
}  // namespace

#define SHAPE(Sys, S_, I_) \
  if (S == S_ && I == I_) return launch_one<Sys, S_, I_>(P, has_loc, c, thr);
#define ALL_SHAPES(Sys) \
  SHAPE(Sys, 7, 5) SHAPE(Sys, 5, 4) SHAPE(Sys, 4, 2) SHAPE(Sys, 3, 2) SHAPE(Sys, 2, 2) SHAPE(Sys, 1, 2)

#define DFB_CAT2(a, b) a##b
#define DFB_CAT(a, b) DFB_CAT2(a, b)

// returns 0 = launched, >0 = CUDA error, -1 = (system, shape) not compiled in this group
int DFB_CAT(launch_general_g, DFB_GROUP)(const DevProblem& P, int system_id, bool has_loc,
                                         const LaunchCfg& c, int thr) {
  const int S = P.rk.S, I = P.rk.I;
#ifdef DFB_PLUGIN_SYSTEM
  (void)system_id;
  SHAPE(DFB_PLUGIN_SYSTEM, 7, 5) SHAPE(DFB_PLUGIN_SYSTEM, 5, 4) SHAPE(DFB_PLUGIN_SYSTEM, 4, 2)
  return -1;
#else
  switch (system_id) {
#if DFB_GROUP == 1
    case DFB_SYS_LORENZ: ALL_SHAPES(SysLorenz) break;
    case DFB_SYS_LORENZ_VARIATIONAL:
      SHAPE(SysLorenzVariational, 7, 5) SHAPE(SysLorenzVariational, 4, 2) break;
#elif DFB_GROUP == 2
    case DFB_SYS_LINEAR1: ALL_SHAPES(SysLinear1) break;
    case DFB_SYS_LINEAR1_SIN: SHAPE(SysLinear1Sin, 7, 5) break;
    case DFB_SYS_LINEAR3: SHAPE(SysLinear3, 7, 5) SHAPE(SysLinear3, 4, 2) break;
    case DFB_SYS_LINEAR3_MATRIX: SHAPE(SysLinear3Matrix, 7, 5) break;
#elif DFB_GROUP == 3
    case DFB_SYS_HARMONIC: ALL_SHAPES(SysHarmonic) break;
    case DFB_SYS_POWER_DECAY: SHAPE(SysPowerDecay, 7, 5) break;
    case DFB_SYS_CONSTANT3: SHAPE(SysConstant3, 7, 5) SHAPE(SysConstant3, 1, 2) break;
    case DFB_SYS_TIME3: SHAPE(SysTime3, 7, 5) SHAPE(SysTime3, 2, 2) break;
    case DFB_SYS_ZERO1: SHAPE(SysZero1, 7, 5) SHAPE(SysZero1, 4, 2) SHAPE(SysZero1, 1, 2) break;
#elif DFB_GROUP == 4
    case DFB_SYS_DDE_LINEAR: SHAPE(SysDdeLinear, 7, 5) SHAPE(SysDdeLinear, 5, 4) SHAPE(SysDdeLinear, 4, 2) break;
    case DFB_SYS_NDDE_LINEAR: SHAPE(SysNddeLinear, 7, 5) SHAPE(SysNddeLinear, 5, 4) break;
    case DFB_SYS_DDE_RELAY_CLAMP: SHAPE(SysDdeRelayClamp, 7, 5) break;
#elif DFB_GROUP == 5
    case DFB_SYS_BOUNCING_BALL: SHAPE(SysBouncingBall, 7, 5) SHAPE(SysBouncingBall, 5, 4) SHAPE(SysBouncingBall, 4, 2) break;
    case DFB_SYS_EGG_BILLIARD: SHAPE(SysEggBilliard, 7, 5) break;
    case DFB_SYS_RELAY_MSIGN: SHAPE(SysRelayMsign, 7, 5) break;
#endif
    default: break;
  }
  (void)S;
  (void)I;
  return -1;
#endif
}

}  // namespace DFB_NS
