// The Lorenz system (examples/ode-chaos-lorenz.rs:19-22) written as a USER system: must give
// bit-identical results to the built-in SysLorenz in every kernel family.
#include "user_system.cuh"

namespace DFB_NS {

struct SysUserLorenz : SysBase {
  static constexpr int N = 3, NP = 3;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    const double x = s.p[0], y = s.p[1], z = s.p[2];
    out[0] = prm[0] * (y - x);
    out[1] = x * (prm[1] - z) - y;
    out[2] = x * y - prm[2] * z;
  }
};

}  // namespace DFB_NS
