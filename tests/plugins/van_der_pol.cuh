// A user system the library does not ship: Van der Pol oscillator x'' - mu (1 - x^2) x' + x = 0,
// with a detector on x (Poincare section x = 0) and a callback that counts crossings.
#include "user_system.cuh"

namespace DFB_NS {

struct SysVanDerPol : SysBase {
  static constexpr int N = 2, NP = 1, NA = 1, ND = 1, NC = 1;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    out[0] = s.p[1];
    out[1] = prm[0] * (1.0 - s.p[0] * s.p[0]) * s.p[1] - s.p[0];
  }
  template <class V>
  __device__ static double detector(int, const V& s, const double*) { return s.p[0]; }
  template <class M>
  __device__ static void callback(int id, M&, const double*, double* aux) {
    if (id == 1) aux[0] += 1.0;
  }
};

}  // namespace DFB_NS
