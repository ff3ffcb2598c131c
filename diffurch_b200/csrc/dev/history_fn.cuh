// Registered history functions: `InitialCondition::eval::<D>` of the reference
// (src/initial_condition.rs:15-50) for the shipped `InitFn` closures.
#pragma once
#include "common.cuh"

namespace DFB_NS {

// y(t) (D = 0) or y'(t) (D = 1) for t <= t_init.  `y0` is the constant initial
// state, `k` the parameter of the sin histories.  Sets DFB_TRAJ_NO_DERIVATIVE where
// the reference hits `unimplemented!` (InitFn<F, ()> with D >= 1).
template <int D, int N>
__device__ __forceinline__ void init_condition_eval(int history_id, double t, const double* y0,
                                                    double k, uint32_t& status, double* out) {
#pragma unroll
  for (int n = 0; n < N; ++n) out[n] = 0.0;
  switch (history_id) {
    case DFB_HIST_CONSTANT:  // constants: all derivatives are zero (:15-22)
      if (D == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) out[n] = y0[n];
      }
      break;
    case DFB_HIST_SIN:  // InitFn(|t| sin(k t), |t| k cos(k t))
      out[0] = D == 0 ? sin(k * t) : k * cos(k * t);
      break;
    case DFB_HIST_SIN_NODERIV:  // InitFn(|t| sin(k t), ())
      if (D == 0) out[0] = sin(k * t);
      else {
        status |= DFB_TRAJ_NO_DERIVATIVE;
        out[0] = __longlong_as_double(0x7ff8000000000000LL);
      }
      break;
    case DFB_HIST_INV_SQUARE:  // InitFn(|t| t.powi(-2), ())
      if (D == 0) out[0] = 1.0 / (t * t);
      else {
        status |= DFB_TRAJ_NO_DERIVATIVE;
        out[0] = __longlong_as_double(0x7ff8000000000000LL);
      }
      break;
  }
}

}  // namespace DFB_NS
