// Device functors for `Solver::equation`, the event detectors and the
// mutating callbacks of the reference's shipped examples and tests.
//
// Signature contract (reference src/state/state_fn.rs:13-98): an equation or a
// detector sees a read-only view {t, p, d, t_prev, p_prev, p(tau), d(tau)}; a
// callback gets mutable t and p plus stop_integration().  `prm` holds what the
// Rust closure captures by value, `aux` what it captures by `&mut`.
#pragma once
#include "common.cuh"

namespace DFB_NS {

// ---- views ------------------------------------------------------------------
template <int N, class Hist>
struct StateRef {
  double t;
  const double* p;
  const double* d;
  double t_prev;
  const double* p_prev;
  Hist* hist;
  // StateRef::p(t) / d(t): history only, never the in-flight step (state_fn.rs:46-51)
  __device__ __forceinline__ void p_at(double tt, double* out) const {
    hist->template eval<0>(tt, out);
  }
  __device__ __forceinline__ void d_at(double tt, double* out) const {
    hist->template eval<1>(tt, out);
  }
};

template <int N, class Hist>
struct StateRefMut {
  double* t;
  double* p;
  const double* d;
  double t_prev;
  const double* p_prev;
  Hist* hist;
  __device__ __forceinline__ void p_at(double tt, double* out) const {
    hist->template eval<0>(tt, out);
  }
  __device__ __forceinline__ void d_at(double tt, double* out) const {
    hist->template eval<1>(tt, out);
  }
  __device__ __forceinline__ void stop_integration() { *t = __longlong_as_double(0x7ff0000000000000LL); }
};

// ---- helpers ------------------------------------------------------------------
// f64::signum: +-1 by sign bit, NaN stays NaN
__device__ __forceinline__ double dev_signum(double x) {
  if (x != x) return x;
  return (__double2hiint(x) < 0) ? -1.0 : 1.0;
}
__device__ __forceinline__ double dev_clamp(double x, double lo, double hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}

// 3x3 (row-major a) times 3xM column-major: each entry a left fold, the order
// nalgebra's static-size gemv produces.
template <int M>
__device__ __forceinline__ void dev_matmul3(const double (&a)[3][3], const double* b, double* c) {
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double acc = a[i][0] * b[3 * j + 0];
      acc = a[i][1] * b[3 * j + 1] + acc;
      acc = a[i][2] * b[3 * j + 2] + acc;
      c[3 * j + i] = acc;
    }
}

// Householder QR of a column-major 3x3 with the positive-diagonal convention
// (what nalgebra's `qr().unpack()` returns).  m is overwritten by Q.
__device__ inline void dev_householder_qr3(double* m, double* rdiag) {
  double w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = m[i];
  double diag[3];
#pragma unroll
  for (int ic = 0; ic < 3; ++ic) {
    double sq = 0.0;
#pragma unroll
    for (int r = ic; r < 3; ++r) sq += w[3 * ic + r] * w[3 * ic + r];
    const double norm = sqrt(sq);
    const double x0 = w[3 * ic + ic];
    const double modulus = x0 >= 0.0 ? x0 : -x0;
    const double sign = x0 >= 0.0 ? 1.0 : -1.0;
    const double signed_norm = sign * norm;
    const double factor = (sq + modulus * norm) * 2.0;
    w[3 * ic + ic] += signed_norm;
    if (factor != 0.0) {
      const double sf = sqrt(factor);
#pragma unroll
      for (int r = ic; r < 3; ++r) w[3 * ic + r] /= sf;
      diag[ic] = -signed_norm;
      const double sgn = diag[ic] > 0.0 ? 1.0 : (diag[ic] < 0.0 ? -1.0 : 0.0);
#pragma unroll
      for (int c = ic + 1; c < 3; ++c) {
        double dot = 0.0;
#pragma unroll
        for (int r = ic; r < 3; ++r) dot += w[3 * ic + r] * w[3 * c + r];
        const double f = dot * (sgn * -2.0);
#pragma unroll
        for (int r = ic; r < 3; ++r) w[3 * c + r] = f * w[3 * ic + r] + sgn * w[3 * c + r];
      }
    } else {
      diag[ic] = signed_norm;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) rdiag[i] = fabs(diag[i]);
  double q[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    const double sgn = diag[i] > 0.0 ? 1.0 : (diag[i] < 0.0 ? -1.0 : 0.0);
#pragma unroll
    for (int c = i; c < 3; ++c) {
      double dot = 0.0;
#pragma unroll
      for (int r = i; r < 3; ++r) dot += w[3 * i + r] * q[3 * c + r];
      const double f = dot * (sgn * -2.0);
#pragma unroll
      for (int r = i; r < 3; ++r) q[3 * c + r] = f * w[3 * i + r] + sgn * q[3 * c + r];
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) m[i] = q[i];
}

// ---- systems ------------------------------------------------------------------
struct SysBase {
  static constexpr int NP = 0, NA = 0, ND = 0, NC = 0;
  static constexpr bool uses_history = false;
  // rhs depends on (t, p, params) only: not on the stale d, p_prev/t_prev or the history
  static constexpr bool pure_rhs = true;
  // no functor of the system (rhs, detectors, callbacks) reads the view's `d` (state_fn.rs:17)
  static constexpr bool reads_d = false;
  // the system has exactly one constant delay and says so with `const_delay(prm)`: lets the
  // coefficient-ring DDE kernel (dev/dde_fixed.cuh) take it
  static constexpr bool has_const_delay = false;
  // every detector is a pure function of the view's (t, p, d) and of the history p(.) / d(.): it reads
  // neither t_prev nor p_prev.  Lets dde_event_kernel reuse a detector's value at a step end as the
  // next step's eval_prev instead of evaluating it again (state_fn.rs:181-188)
  static constexpr bool pure_detectors = false;
  template <class V>
  __device__ static double detector(int, const V&, const double*) { return 0.0; }
  template <class M>
  __device__ static void callback(int, M&, const double*, double*) {}
};

struct SysLorenz : SysBase {  // examples/ode-chaos-lorenz.rs:19-22
  static constexpr int ID = DFB_SYS_LORENZ, N = 3, NP = 3;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    const double x = s.p[0], y = s.p[1], z = s.p[2];
    out[0] = prm[0] * (y - x);
    out[1] = x * (prm[1] - z) - y;
    out[2] = x * y - prm[2] * z;
  }
};

struct SysLorenzVariational : SysBase {  // tests/lyapunov_exponents.rs:264-296
  static constexpr int ID = DFB_SYS_LORENZ_VARIATIONAL, N = 12, NP = 3, NA = 3, NC = 1;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    const double sigma = prm[0], rho = prm[1], beta = prm[2];
    const double x = s.p[0], y = s.p[1], z = s.p[2];
    out[0] = sigma * (y - x);
    out[1] = x * (rho - z) - y;
    out[2] = x * y - beta * z;
#if DFB_ARITH == 1
    // FAST: the Jacobian product without its structural zero and unit entries (the reference's dense
    // 3x3 gemm multiplies by 0.0 and -1.0): 11 flop per column instead of 15
    const double rz = rho - z;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double v0 = s.p[3 + 3 * j], v1 = s.p[4 + 3 * j], v2 = s.p[5 + 3 * j];
      out[3 + 3 * j] = sigma * (v1 - v0);
      out[4 + 3 * j] = fma(rz, v0, fma(-x, v2, -v1));
      out[5 + 3 * j] = fma(y, v0, fma(x, v1, -beta * v2));
    }
#else
    const double df[3][3] = {{-sigma, sigma, 0.0}, {rho - z, -1.0, -x}, {y, x, -beta}};
    dev_matmul3<3>(df, s.p + 3, out + 3);
#endif
  }
  template <class M>
  __device__ static void callback(int id, M& s, const double*, double* aux) {
    if (id != 1) return;
    double r[3];
    dev_householder_qr3(s.p + 3, r);
#pragma unroll
    for (int i = 0; i < 3; ++i) aux[i] += log(fabs(r[i]));
  }
};

struct SysLinear1 : SysBase {  // tests/lyapunov_exponents.rs:17-31
  static constexpr int ID = DFB_SYS_LINEAR1, N = 1, NP = 1, NA = 1, NC = 1;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    out[0] = s.p[0] * prm[0];
  }
  template <class M>
  __device__ static void callback(int id, M& s, const double*, double* aux) {
    if (id != 1) return;
    const double norm = fabs(s.p[0]);
    s.p[0] /= norm;
    aux[0] += log(norm);
  }
};

struct SysLinear1Sin : SysBase {  // tests/lyapunov_exponents.rs:55
  static constexpr int ID = DFB_SYS_LINEAR1_SIN, N = 1, NP = 1, NA = 1, NC = 1;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    out[0] = (1.0 + sin(s.t)) * prm[0] * s.p[0];
  }
  template <class M>
  __device__ static void callback(int id, M& s, const double* prm, double* aux) {
    SysLinear1::callback(id, s, prm, aux);
  }
};

struct SysLinear3 : SysBase {  // examples/ode-linear-3-fibonacci.rs:11-20
  static constexpr int ID = DFB_SYS_LINEAR3, N = 3, NP = 9;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    const double a[3][3] = {{prm[0], prm[1], prm[2]}, {prm[3], prm[4], prm[5]}, {prm[6], prm[7], prm[8]}};
    dev_matmul3<1>(a, s.p, out);
  }
};

struct SysLinear3Matrix : SysBase {  // tests/lyapunov_exponents.rs:107-131, 187-210
  static constexpr int ID = DFB_SYS_LINEAR3_MATRIX, N = 9, NP = 9, NA = 3, NC = 2;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    const double a[3][3] = {{prm[0], prm[1], prm[2]}, {prm[3], prm[4], prm[5]}, {prm[6], prm[7], prm[8]}};
    dev_matmul3<3>(a, s.p, out);
  }
  template <class M>
  __device__ static void callback(int id, M& s, const double*, double* aux) {
    if (id != 1 && id != 2) return;
    double r[3];
    dev_householder_qr3(s.p, r);
#pragma unroll
    for (int i = 0; i < 3; ++i) aux[i] += log(r[i]);
  }
};

struct SysHarmonic : SysBase {  // tests/equation_types.rs:89-92
  static constexpr int ID = DFB_SYS_HARMONIC, N = 2;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double*, double* out) {
    out[0] = s.p[1];
    out[1] = -s.p[0];
  }
};

struct SysPowerDecay : SysBase {  // tests/equation_types.rs:107
  static constexpr int ID = DFB_SYS_POWER_DECAY, N = 1;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double*, double* out) {
    out[0] = -2.0 * s.p[0] / s.t;
  }
};

struct SysConstant3 : SysBase {  // tests/equation_types.rs:14
  static constexpr int ID = DFB_SYS_CONSTANT3, N = 3;
  template <class V>
  __device__ __forceinline__ static void rhs(const V&, const double*, double* out) {
    out[0] = 1.0;
    out[1] = -1.0;
    out[2] = 0.0;
  }
};

struct SysTime3 : SysBase {  // tests/equation_types.rs:31
  static constexpr int ID = DFB_SYS_TIME3, N = 3;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double*, double* out) {
    out[0] = 0.0;
    out[1] = 1.0;
    out[2] = s.t;
  }
};

struct SysZero1 : SysBase {  // tests/delay_propagation.rs:13, :86
  static constexpr int ID = DFB_SYS_ZERO1, N = 1, ND = 1;
  template <class V>
  __device__ __forceinline__ static void rhs(const V&, const double*, double* out) {
    out[0] = 0.0;
  }
  template <class V>
  __device__ static double detector(int id, const V& s, const double*) {
    return id == 1 ? s.t / 2.0 : 0.0;
  }
};

struct SysDdeLinear : SysBase {  // tests/equation_types.rs:131
  static constexpr int ID = DFB_SYS_DDE_LINEAR, N = 1, NP = 4;
  static constexpr bool uses_history = true;
  static constexpr bool pure_rhs = false;
  // the one constant delay of this equation (lets the streaming-window kernel take it)
  static constexpr bool has_const_delay = true;
  __device__ __forceinline__ static double const_delay(const double* prm) { return prm[2]; }
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    double del[1];
    s.p_at(s.t - prm[2], del);
    out[0] = prm[0] * s.p[0] + prm[1] * del[0];
  }
};

struct SysNddeLinear : SysBase {  // tests/equation_types.rs:150
  static constexpr int ID = DFB_SYS_NDDE_LINEAR, N = 1, NP = 4;
  static constexpr bool uses_history = true;
  static constexpr bool pure_rhs = false;
  static constexpr bool has_const_delay = true;
  __device__ __forceinline__ static double const_delay(const double* prm) { return prm[2]; }
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    double del[1], ddel[1];
    s.p_at(s.t - prm[2], del);
    s.d_at(s.t - prm[2], ddel);
    out[0] = prm[0] * del[0] + prm[1] * ddel[0];
  }
};

struct SysBouncingBall : SysBase {  // examples/bouncing-ball.rs:21-33
  static constexpr int ID = DFB_SYS_BOUNCING_BALL, N = 2, NP = 2, ND = 2, NC = 1;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    out[0] = s.p[1];
    out[1] = -prm[0];
  }
  template <class V>
  __device__ static double detector(int id, const V& s, const double*) {
    return id == 1 ? s.p[1] : s.p[0];
  }
  template <class M>
  __device__ static void callback(int id, M& s, const double* prm, double*) {
    if (id != 1) return;
    s.p[0] = 0.0;
    s.p[1] = prm[1] * fabs(s.p[1]);
  }
};

struct SysEggBilliard : SysBase {  // examples/egg-billiard.rs:12-40
  static constexpr int ID = DFB_SYS_EGG_BILLIARD, N = 4, NP = 1, NA = 1, ND = 1, NC = 1;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double*, double* out) {
    out[0] = s.p[2];
    out[1] = s.p[3];
    out[2] = 0.0;
    out[3] = 0.0;
  }
  template <class V>
  __device__ static double detector(int, const V& s, const double*) {
    const double x = s.p[0], y = s.p[1];
    // powi(2) = y*y, powi(3) = y*(y*y)
#if DFB_ARITH == 1
    // FAST: the division by the constant 3 (evaluated ~45 times per located event) as a multiply
    return x * x + y * y - (y * (y * y)) * (1.0 / 3.0) - 1.0;
#else
    return x * x + y * y - (y * (y * y)) / 3.0 - 1.0;
#endif
  }
  template <class M>
  __device__ static void callback(int id, M& s, const double* prm, double* aux) {
    if (id != 1) return;
    double* v = s.p;
    const double x_normal = 2.0 * v[0];
    const double y_normal = 2.0 * v[1] - v[1] * v[1];
    const double h = x_normal * x_normal + y_normal * y_normal;
    const double k1 = -(x_normal * x_normal - y_normal * y_normal) / h;
    const double k2 = -2.0 * x_normal * y_normal / h;
    const double n2 = k1 * v[2] + k2 * v[3];
    const double n3 = k2 * v[2] - k1 * v[3];
    v[2] = n2;
    v[3] = n3;
    aux[0] += 1.0;
    if (aux[0] >= prm[0]) s.stop_integration();
  }
};

struct SysRelayMsign : SysBase {  // examples/ode-2-relay-msign.rs:17-20
  static constexpr int ID = DFB_SYS_RELAY_MSIGN, N = 2, ND = 1;
  static constexpr bool pure_rhs = false;  // reads p_prev
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double*, double* out) {
    out[0] = s.p[1];
    out[1] = -2.0 * dev_signum(s.p_prev[0]);
  }
  template <class V>
  __device__ static double detector(int, const V& s, const double*) { return s.p[0]; }
};

struct SysDdeRelayClamp : SysBase {  // examples/dde-relay-clamp.rs:24-35
  static constexpr int ID = DFB_SYS_DDE_RELAY_CLAMP, N = 2, NP = 3, ND = 1;
  static constexpr bool uses_history = true;
  static constexpr bool pure_rhs = false;
  static constexpr bool pure_detectors = true;
  static constexpr bool has_const_delay = true;  // rhs looks the history up at s.t - const_delay(prm) only
  __device__ __forceinline__ static double const_delay(const double* prm) { return prm[2]; }
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    const double x = s.p[0], dx = s.p[1];
    double del[2];
    s.p_at(s.t - prm[2], del);
    out[0] = dx;
    out[1] = -prm[1] * dx - x - prm[0] * dev_clamp(del[0], -1.0, 1.0);
  }
  template <class V>
  __device__ static double detector(int, const V& s, const double* prm) {
    double del[2];
    s.p_at(s.t - prm[2], del);
    return fabs(del[0]) - 1.0;
  }
};

}  // namespace DFB_NS

// A user-system plugin build (diffurch_b200.user.build_user_system) compiles the kernel
// translation units with -DDFB_PLUGIN_HEADER="<user header>" -DDFB_PLUGIN_SYSTEM=<struct>: the
// header is pulled in here, after the views and SysBase it builds on.
#ifdef DFB_PLUGIN_HEADER
#include DFB_PLUGIN_HEADER
#endif
