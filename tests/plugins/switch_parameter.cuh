// examples/loc-switch-parameter.rs (old-API example of the reference, commented out upstream): a
// bouncing particle whose gravity flips sign at every crossing of x = 0.  Upstream the parameter
// lives in a `Cell` captured by both the equation and the callback; here a callback may write the
// per-trajectory parameter block (`double* prm`), which the equation reads on the next evaluation.
#include "user_system.cuh"

namespace DFB_NS {

struct SysSwitchParameter : SysBase {
  static constexpr int N = 2, NP = 2, NA = 1, ND = 2, NC = 1;  // prm = {g, k}; aux[0] counts switchings
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    out[0] = s.p[1];
    out[1] = -prm[0];
  }
  template <class V>
  __device__ static double detector(int id, const V& s, const double*) { return id == 1 ? s.p[1] : s.p[0]; }
  template <class M>
  __device__ static void callback(int id, M& s, double* prm, double* aux) {
    if (id != 1) return;
    s.p[1] *= prm[1];   // *dx *= k
    prm[0] = -prm[0];   // g.set(g.get() * -1.)
    aux[0] += 1.0;
  }
};

}  // namespace DFB_NS
