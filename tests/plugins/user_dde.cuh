// x'(t) = a x(t) + b x(t - tau) (tests/equation_types.rs:131) written as a USER delay system:
// takes the coefficient-ring DDE kernel and the general kernel like the built-in SysDdeLinear.
#include "user_system.cuh"

namespace DFB_NS {

struct SysUserDde : SysBase {
  static constexpr int N = 1, NP = 4;  // {a, b, tau, k}; k = parameter of the sin history (params[3])
  static constexpr bool uses_history = true;
  static constexpr bool pure_rhs = false;
  static constexpr bool has_const_delay = true;
  __device__ __forceinline__ static double const_delay(const double* prm) { return prm[2]; }
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    double del[1];
    s.p_at(s.t - prm[2], del);
    out[0] = prm[0] * s.p[0] + prm[1] * del[0];
  }
};

}  // namespace DFB_NS
