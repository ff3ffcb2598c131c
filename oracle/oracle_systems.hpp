// =============================================================================
// oracle_systems.hpp — TEST INFRASTRUCTURE ONLY (see diffurch_oracle.hpp).
//
// The closures of the reference's live examples and tests, restated against
// the oracle's StateRef / StateRefMut, keyed by dfb_system_id.  Parameters are
// what the Rust closures capture by value; `aux` is what they capture by
// `&mut` (Lyapunov log sums, reflection counter).
// =============================================================================
#pragma once
#include <cmath>

#include "../include/diffurch_b200.h"
#include "diffurch_oracle.hpp"

namespace orc {

// f64::signum (core): 1.0 for +0.0 and positives, -1.0 for -0.0 and negatives, NaN for NaN.
inline double rust_signum(double x) {
  if (std::isnan(x)) return x;
  return std::signbit(x) ? -1.0 : 1.0;
}
inline double rust_clamp(double x, double lo, double hi) {
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}

// ---------------------------------------------------------------------------
// nalgebra 0.34.1 (Cargo.lock:764-767, source NOT under /root/reference) —
// restated from its published algorithms:
//   * static-size `A * B` = column-by-column gemv, each entry a left fold
//     ((a_i0*b_0j) + a_i1*b_1j) + a_i2*b_2j  (base/blas.rs gemv/axcpy).
//   * `qr().unpack()` = Householder reflections (linalg/householder.rs),
//     R returned with the MODULUS of the diagonal, Q accumulated with the
//     matching signs, i.e. the unique QR with a positive diagonal.
// Rounding-level details are pinned only through the Lyapunov tolerances of
// tests/lyapunov_exponents.rs:161,237,309.
// ---------------------------------------------------------------------------
// C(3xM, column-major) = A(3x3 row-major entries a[r][c]) * B(3xM column-major)
inline void matmul3(const double a[3][3], const double* b, double* c, int m) {
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < 3; ++i) {
      double acc = a[i][0] * b[3 * j + 0];
      acc = a[i][1] * b[3 * j + 1] + acc;
      acc = a[i][2] * b[3 * j + 2] + acc;
      c[3 * j + i] = acc;
    }
}

// In-place Householder QR of a column-major 3x3; q receives Q (column-major),
// rdiag the (positive) diagonal of R.
inline void householder_qr3(const double* a_in, double* q, double* rdiag) {
  double m[9];
  for (int i = 0; i < 9; ++i) m[i] = a_in[i];
  double diag[3];
  auto at = [&](int r, int c) -> double& { return m[3 * c + r]; };
  for (int ic = 0; ic < 3; ++ic) {
    // reflection_axis_mut on rows ic.. of column ic
    double sq = 0.0;
    for (int r = ic; r < 3; ++r) sq += at(r, ic) * at(r, ic);
    double norm = std::sqrt(sq);
    double x0 = at(ic, ic);
    double modulus = x0 >= 0.0 ? x0 : -x0;
    double sign = x0 >= 0.0 ? 1.0 : -1.0;
    double signed_norm = sign * norm;
    double factor = (sq + modulus * norm) * 2.0;
    at(ic, ic) += signed_norm;
    if (factor != 0.0) {
      double sf = std::sqrt(factor);
      for (int r = ic; r < 3; ++r) at(r, ic) /= sf;
      diag[ic] = -signed_norm;
      // reflect the trailing columns: col = factor*axis + sgn*col
      double sgn = diag[ic] > 0.0 ? 1.0 : (diag[ic] < 0.0 ? -1.0 : 0.0);
      for (int c = ic + 1; c < 3; ++c) {
        double dot = 0.0;
        for (int r = ic; r < 3; ++r) dot += at(r, ic) * at(r, c);
        double m_two = sgn * -2.0;
        double f = dot * m_two;
        for (int r = ic; r < 3; ++r) at(r, c) = f * at(r, ic) + sgn * at(r, c);
      }
    } else {
      diag[ic] = signed_norm;
    }
  }
  for (int i = 0; i < 3; ++i) rdiag[i] = std::fabs(diag[i]);
  // Q = H_0 H_1 H_2 applied to the identity, last reflection first
  for (int i = 0; i < 9; ++i) q[i] = 0.0;
  q[0] = q[4] = q[8] = 1.0;
  auto qa = [&](int r, int c) -> double& { return q[3 * c + r]; };
  for (int i = 2; i >= 0; --i) {
    double sgn = diag[i] > 0.0 ? 1.0 : (diag[i] < 0.0 ? -1.0 : 0.0);
    for (int c = i; c < 3; ++c) {
      double dot = 0.0;
      for (int r = i; r < 3; ++r) dot += at(r, i) * qa(r, c);
      double m_two = sgn * -2.0;
      double f = dot * m_two;
      for (int r = i; r < 3; ++r) qa(r, c) = f * at(r, i) + sgn * qa(r, c);
    }
  }
}

// ---------------------------------------------------------------------------
// Systems.  Each exposes: N, rhs(prm) -> callable(StateRef) -> Vec<N>,
// detector(id, prm), callback(id, prm, aux).
// ---------------------------------------------------------------------------
template <int N_, int S, int I>
struct SysBase {
  static constexpr int N = N_;
  using Ref = StateRef<N_, S, I>;
  using RefMut = StateRefMut<N_, S, I>;
  using Det = std::function<double(const Ref&)>;
  using Cb = std::function<void(RefMut&)>;
  static Det detector(int, const double*) { return Det(); }
  static Cb callback(int, const double*, double*) { return Cb(); }
};

template <int S, int I>
struct Lorenz : SysBase<3, S, I> {  // examples/ode-chaos-lorenz.rs:19-22
  using Ref = StateRef<3, S, I>;
  static auto rhs(const double* prm) {
    const double sigma = prm[0], rho = prm[1], beta = prm[2];
    return [=](const Ref& s) {
      const double x = s.p[0], y = s.p[1], z = s.p[2];
      return Vec<3>{{sigma * (y - x), x * (rho - z) - y, x * y - beta * z}};
    };
  }
};

template <int S, int I>
struct LorenzVariational : SysBase<12, S, I> {  // tests/lyapunov_exponents.rs:264-296
  using Ref = StateRef<12, S, I>;
  using RefMut = StateRefMut<12, S, I>;
  using Cb = std::function<void(RefMut&)>;
  static auto rhs(const double* prm) {
    const double sigma = prm[0], rho = prm[1], beta = prm[2];
    return [=](const Ref& s) {
      const double x = s.p[0], y = s.p[1], z = s.p[2];
      Vec<12> diff = Vec<12>::zero();
      diff[0] = sigma * (y - x);
      diff[1] = x * (rho - z) - y;
      diff[2] = x * y - beta * z;
      const double df[3][3] = {{-sigma, sigma, 0.0}, {rho - z, -1.0, -x}, {y, x, -beta}};
      matmul3(df, &s.p.v[3], &diff.v[3], 3);
      return diff;
    };
  }
  static Cb callback(int id, const double*, double* aux) {
    if (id != 1) return Cb();
    return [aux](RefMut& s) {  // :289-296
      double q[9], r[3];
      householder_qr3(&s.p.v[3], q, r);
      for (int i = 0; i < 9; ++i) s.p.v[3 + i] = q[i];
      for (int i = 0; i < 3; ++i) aux[i] += std::log(std::fabs(r[i]));
    };
  }
};

template <int S, int I>
struct Linear1 : SysBase<1, S, I> {  // tests/lyapunov_exponents.rs:17, equation_types.rs:76
  using Ref = StateRef<1, S, I>;
  using RefMut = StateRefMut<1, S, I>;
  using Cb = std::function<void(RefMut&)>;
  static auto rhs(const double* prm) {
    const double lambda = prm[0];
    return [=](const Ref& s) { return Vec<1>{{s.p[0] * lambda}}; };
  }
  static Cb callback(int id, const double*, double* aux) {
    if (id != 1) return Cb();
    return [aux](RefMut& s) {  // :26-30
      double norm = std::fabs(s.p[0]);
      s.p[0] /= norm;
      aux[0] += std::log(norm);
    };
  }
};

template <int S, int I>
struct Linear1Sin : SysBase<1, S, I> {  // tests/lyapunov_exponents.rs:55
  using Ref = StateRef<1, S, I>;
  using RefMut = StateRefMut<1, S, I>;
  using Cb = std::function<void(RefMut&)>;
  static auto rhs(const double* prm) {
    const double lambda = prm[0];
    return [=](const Ref& s) { return Vec<1>{{(1.0 + std::sin(s.t)) * lambda * s.p[0]}}; };
  }
  static Cb callback(int id, const double* prm, double* aux) {
    return Linear1<S, I>::callback(id, prm, aux);
  }
};

template <int S, int I>
struct Linear3 : SysBase<3, S, I> {  // examples/ode-linear-3-fibonacci.rs:11-20
  using Ref = StateRef<3, S, I>;
  static auto rhs(const double* prm) {
    double a[3][3];
    for (int i = 0; i < 9; ++i) a[i / 3][i % 3] = prm[i];
    return [=](const Ref& s) {
      Vec<3> r;
      matmul3(a, s.p.v, r.v, 1);
      return r;
    };
  }
};

template <int S, int I>
struct Linear3Matrix : SysBase<9, S, I> {  // tests/lyapunov_exponents.rs:107-131, 187-210
  using Ref = StateRef<9, S, I>;
  using RefMut = StateRefMut<9, S, I>;
  using Cb = std::function<void(RefMut&)>;
  static auto rhs(const double* prm) {
    double a[3][3];
    for (int i = 0; i < 9; ++i) a[i / 3][i % 3] = prm[i];
    return [=](const Ref& s) {
      Vec<9> r;
      matmul3(a, s.p.v, r.v, 3);
      return r;
    };
  }
  static Cb callback(int id, const double*, double* aux) {
    if (id != 1 && id != 2) return Cb();
    return [aux](RefMut& s) {
      // id 1 (:127) takes ln of the diagonal nalgebra returns, which is its
      // modulus, so ids 1 and 2 (:206, ln|r|) compute the same thing.
      double q[9], r[3];
      householder_qr3(s.p.v, q, r);
      for (int i = 0; i < 9; ++i) s.p.v[i] = q[i];
      for (int i = 0; i < 3; ++i) aux[i] += std::log(r[i]);
    };
  }
};

template <int S, int I>
struct Harmonic : SysBase<2, S, I> {  // tests/equation_types.rs:89-92
  using Ref = StateRef<2, S, I>;
  static auto rhs(const double*) {
    return [](const Ref& s) { return Vec<2>{{s.p[1], -s.p[0]}}; };
  }
};

template <int S, int I>
struct PowerDecay : SysBase<1, S, I> {  // tests/equation_types.rs:107
  using Ref = StateRef<1, S, I>;
  static auto rhs(const double*) {
    return [](const Ref& s) { return Vec<1>{{-2.0 * s.p[0] / s.t}}; };
  }
};

template <int S, int I>
struct Constant3 : SysBase<3, S, I> {  // tests/equation_types.rs:14
  using Ref = StateRef<3, S, I>;
  static auto rhs(const double*) {
    return [](const Ref&) { return Vec<3>{{1.0, -1.0, 0.0}}; };
  }
};

template <int S, int I>
struct Time3 : SysBase<3, S, I> {  // tests/equation_types.rs:31
  using Ref = StateRef<3, S, I>;
  static auto rhs(const double*) {
    return [](const Ref& s) { return Vec<3>{{0.0, 1.0, s.t}}; };
  }
};

template <int S, int I>
struct Zero1 : SysBase<1, S, I> {  // tests/delay_propagation.rs:13
  using Ref = StateRef<1, S, I>;
  using Det = std::function<double(const Ref&)>;
  static auto rhs(const double*) {
    return [](const Ref&) { return Vec<1>{{0.0}}; };
  }
  static Det detector(int id, const double*) {
    if (id == 1) return [](const Ref& s) { return s.t / 2.0; };  // pantograph :86
    return Det();
  }
};

template <int S, int I>
struct DdeLinear : SysBase<1, S, I> {  // tests/equation_types.rs:131
  using Ref = StateRef<1, S, I>;
  static auto rhs(const double* prm) {
    const double a = prm[0], b = prm[1], tau = prm[2];
    return [=](const Ref& s) { return Vec<1>{{a * s.p[0] + b * s.p_at(s.t - tau)[0]}}; };
  }
};

template <int S, int I>
struct NddeLinear : SysBase<1, S, I> {  // tests/equation_types.rs:150
  using Ref = StateRef<1, S, I>;
  static auto rhs(const double* prm) {
    const double a = prm[0], b = prm[1], tau = prm[2];
    return [=](const Ref& s) {
      return Vec<1>{{a * s.p_at(s.t - tau)[0] + b * s.d_at(s.t - tau)[0]}};
    };
  }
};

template <int S, int I>
struct BouncingBall : SysBase<2, S, I> {  // examples/bouncing-ball.rs:21-33
  using Ref = StateRef<2, S, I>;
  using RefMut = StateRefMut<2, S, I>;
  using Det = std::function<double(const Ref&)>;
  using Cb = std::function<void(RefMut&)>;
  static auto rhs(const double* prm) {
    const double g = prm[0];
    return [=](const Ref& s) { return Vec<2>{{s.p[1], -g}}; };
  }
  static Det detector(int id, const double*) {
    if (id == 1) return [](const Ref& s) { return s.p[1]; };  // :27
    if (id == 2) return [](const Ref& s) { return s.p[0]; };  // :28
    return Det();
  }
  static Cb callback(int id, const double* prm, double*) {
    if (id != 1) return Cb();
    const double k = prm[1];
    return [k](RefMut& s) {  // :28-31
      s.p[0] = 0.0;
      s.p[1] = k * std::fabs(s.p[1]);
    };
  }
};

template <int S, int I>
struct EggBilliard : SysBase<4, S, I> {  // examples/egg-billiard.rs:12-40
  using Ref = StateRef<4, S, I>;
  using RefMut = StateRefMut<4, S, I>;
  using Det = std::function<double(const Ref&)>;
  using Cb = std::function<void(RefMut&)>;
  static auto rhs(const double*) {
    return [](const Ref& s) { return Vec<4>{{s.p[2], s.p[3], 0.0, 0.0}}; };
  }
  static Det detector(int id, const double*) {
    if (id != 1) return Det();
    return [](const Ref& s) {  // :18-21
      const double x = s.p[0], y = s.p[1];
      return powi(x, 2) + powi(y, 2) - powi(y, 3) / 3.0 - 1.0;
    };
  }
  static Cb callback(int id, const double* prm, double* aux) {
    if (id != 1) return Cb();
    const double max_reflections = prm[0];
    return [aux, max_reflections](RefMut& s) {  // :22-38
      Vec<4>& v = s.p;
      double x_normal = 2.0 * v[0];
      double y_normal = 2.0 * v[1] - powi(v[1], 2);
      double h = powi(x_normal, 2) + powi(y_normal, 2);
      double k1 = -(powi(x_normal, 2) - powi(y_normal, 2)) / h;
      double k2 = -2.0 * x_normal * y_normal / h;
      double n2 = k1 * v[2] + k2 * v[3];
      double n3 = k2 * v[2] - k1 * v[3];
      v[2] = n2;
      v[3] = n3;
      aux[0] += 1.0;
      if (aux[0] >= max_reflections) s.stop_integration();
    };
  }
};

template <int S, int I>
struct RelayMsign : SysBase<2, S, I> {  // examples/ode-2-relay-msign.rs:17-20
  using Ref = StateRef<2, S, I>;
  using Det = std::function<double(const Ref&)>;
  static auto rhs(const double*) {
    return [](const Ref& s) { return Vec<2>{{s.p[1], -2.0 * rust_signum(s.p_prev[0])}}; };
  }
  static Det detector(int id, const double*) {
    if (id == 1) return [](const Ref& s) { return s.p[0]; };
    return Det();
  }
};

template <int S, int I>
struct DdeRelayClamp : SysBase<2, S, I> {  // examples/dde-relay-clamp.rs:24-35
  using Ref = StateRef<2, S, I>;
  using Det = std::function<double(const Ref&)>;
  static auto rhs(const double* prm) {
    const double alpha = prm[0], sigma = prm[1], T = prm[2];
    return [=](const Ref& s) {
      const double x = s.p[0], dx = s.p[1];
      return Vec<2>{{dx, -sigma * dx - x - alpha * rust_clamp(s.p_at(s.t - T)[0], -1.0, 1.0)}};
    };
  }
  static Det detector(int id, const double* prm) {
    if (id != 1) return Det();
    const double T = prm[2];
    return [T](const Ref& s) { return std::fabs(s.p_at(s.t - T)[0]) - 1.0; };
  }
};

}  // namespace orc
