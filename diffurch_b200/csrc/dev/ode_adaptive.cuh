// Hot path #1b: adaptive-step explicit RK over an ODE ensemble, no events, no delay
// history, FAST arithmetic (reference: Solver::run solver.rs:292-310 with an
// `AutomaticStepsize` controller, stepsize.rs:35-85; State::make_step state.rs:94-120).
//
// Like the fixed-step kernel every trajectory lives in registers, so the bound is the
// FP64 pipe.  What is different is that trajectories take different numbers of steps:
//
//  * every lane carries its own (t, h); an attempted step is the same instruction stream
//    for every lane whether it will be accepted or rejected, so rejections never diverge
//    the warp (the reference's `undo_step` + new `make_step` is simply "do not commit");
//  * lanes that finish early do not idle until the slowest lane of their warp is done:
//    a finished lane stores its result and pulls the next trajectory index from a global
//    work counter (dynamic trajectory assignment; results are indexed by trajectory, so
//    they do not depend on which lane integrated what);
//  * the controller's `(1/err).powf(1/(order+1))` — one libm `pow` per attempted step in
//    the reference — is a MUFU-seeded Newton step on f^-p = err (relative error
//    ~1e-11 on a quantity that only scales the next step size), and its N divisions use a
//    MUFU-seeded reciprocal; both fall back to the IEEE `pow` / division outside the range
//    where the shortcut is safe, so the corner cases (err = 0, inf, NaN) behave exactly as
//    the reference's formula does.  Accept/reject is decided on err itself, never on the root.
//
// Only systems whose right-hand side is a pure function of (t, p, params) may use this
// kernel (`Sys::pure_rhs`): the reference re-evaluates k[0] after an `undo_step`
// (state.rs:97-98), which for a pure function reproduces the value already held.
#pragma once
#include "common.cuh"
#include "ode_fixed.cuh"
#include "systems.cuh"

namespace DFB_NS {

struct OdeAdaptiveArgs {
  double a[DFB_MAX_STAGES * DFB_MAX_STAGES];
  double b[DFB_MAX_STAGES];
  double c[DFB_MAX_STAGES];
  double e[DFB_MAX_STAGES];  // b2_j - b_j (host-computed, one IEEE subtraction as in state.rs:118)
  dfb_auto_stepsize ctl;
  double t0, t1;
  double inv_p;   // 1 / (order + 1) as the reference forms it (stepsize.rs:74)
  float inv_pf;
  int32_t p;      // order + 1
  size_t n_traj;
  const double* y0;
  const double* params;
  double* t_final;
  double* y_final;
  unsigned long long* steps;
  unsigned long long* rejected;
  unsigned long long* rhs_evals;
  uint32_t* status;
  unsigned long long* queue;  // next unassigned trajectory (zeroed before the launch)
  // Trace output policy (on_step closures, solver.rs:290, 309); read by the TRACE instantiations
  int32_t trace_every, trace_capacity;
  double* trace_t;      // [cap][n]
  double* trace_y;      // [cap][N][n]
  uint32_t* n_samples;  // [n]
};

// 1/x for a positive, normal x: MUFU.RCP64H seed + two Newton steps (relative error ~1e-16).
__device__ __forceinline__ double fast_rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
// exponent field strictly inside (1, 2045): positive-or-negative normal number far from both ends
__device__ __forceinline__ bool safely_normal(double x) {
  const unsigned hi = unsigned(__double2hiint(x)) & 0x7fffffffu;
  return (hi - 0x06400000u) < 0x73000000u;  // 2^-923 .. 2^+923
}

// f^P by repeated multiplication (P compile-time) or a loop (P = 0: run-time exponent p)
template <int P>
__device__ __forceinline__ double pow_int(double f, int p) {
  if (P == 5) {
    const double f2 = f * f;
    return f2 * f2 * f;
  }
  if (P > 0) {
    double r = f;
#pragma unroll
    for (int q = 1; q < P; ++q) r *= f;
    return r;
  }
  double r = f;
  for (int q = 1; q < p; ++q) r *= f;
  return r;
}

// (1/err)^(1/p) for a finite err in [2^-99, 2^99): MUFU.LG2/EX2 seed (~1e-6 relative) + one Newton
// step on f^-p = err, f <- f ((p+1) - err f^p) / p  (relative error ~1e-11: it scales a step size,
// it is never compared against anything).
template <int P>
__device__ __forceinline__ double inv_root_fast(double err, int p, double inv_p, float inv_pf) {
  float lg, g;
  const float ef = float(err);
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(ef));
  const float x = -inv_pf * lg;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(x));
  const double f = double(g);
  const double pp1 = double((P > 0 ? P : p) + 1);
  return f * fma(-err, pow_int<P>(f, p), pp1) * inv_p;
}
// non-negative finite x in [2^-99, 2^99)
__device__ __forceinline__ bool in_root_range(double x) {
  return (unsigned(__double2hiint(x)) - 0x39c00000u) < (0x46200000u - 0x39c00000u);
}

// P = order + 1 when it is one of the compiled exponents, else 0 (run-time exponent A.p)
template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, unsigned EMASK, int TPT, int P, bool TRACE = false>
struct OdeAdaptiveStepper {
  static constexpr int N = Sys::N;
  static constexpr int NP = Sys::NP > 0 ? Sys::NP : 1;
  using Ref = StateRef<N, NoHist>;

  const OdeAdaptiveArgs& A;
  size_t idx[TPT];
  bool active[TPT];
  double t[TPT], h_auto[TPT];
  double p[TPT][N], d[TPT][N];
  double prm[TPT][NP];
  unsigned n_steps[TPT], n_rej[TPT];  // per trajectory; 2^32 attempts of one member is beyond any run
  uint32_t flags[TPT];
  unsigned n_samp[TPT];
  NoHist nohist;

  __device__ explicit OdeAdaptiveStepper(const OdeAdaptiveArgs& A_) : A(A_) {}

  __device__ __forceinline__ void rhs(int u, double tt, const double* y, double* out) {
    Sys::rhs(Ref{tt, y, d[u], tt, y, &nohist}, prm[u], out);
  }

  // on_step: the n_steps-th invocation (0 = the initial state) is stored when n_steps % every == 0.
  // Every member has its own time grid, so the rows are member-indexed, not warp-uniform.
  __device__ __forceinline__ void sample(int u) {
    if (!TRACE) return;
    if (active[u] && n_steps[u] % unsigned(A.trace_every) == 0u && n_samp[u] < unsigned(A.trace_capacity)) {
      const size_t n_traj = A.n_traj, i = idx[u];
      A.trace_t[size_t(n_samp[u]) * n_traj + i] = t[u];
#pragma unroll
      for (int n = 0; n < N; ++n) A.trace_y[(size_t(n_samp[u]) * N + n) * n_traj + i] = p[u][n];
      ++n_samp[u];
    }
  }

  __device__ void fetch(int u) {
    const unsigned long long i = atomicAdd(A.queue, 1ull);
    active[u] = i < A.n_traj;
    if (!active[u]) return;
    idx[u] = size_t(i);
    const size_t n_traj = A.n_traj;
#pragma unroll
    for (int n = 0; n < Sys::NP; ++n) prm[u][n] = A.params[size_t(n) * n_traj + i];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      p[u][n] = A.y0[size_t(n) * n_traj + i];
      d[u][n] = 0.0;
    }
    t[u] = A.t0;
    h_auto[u] = A.ctl.stepsize;
    n_steps[u] = n_rej[u] = 0u;
    flags[u] = 0u;
    n_samp[u] = 0u;
    sample(u);  // events_on_step at t_init (solver.rs:290)
    if (t[u] < A.t1) rhs(u, t[u], p[u], d[u]);  // k[0] of the first make_step (state.rs:97-98)
  }

  __device__ void store(int u) {
    const size_t n_traj = A.n_traj, i = idx[u];
    bool finite = true;
#pragma unroll
    for (int n = 0; n < N; ++n) finite = finite && (fabs(p[u][n]) < __longlong_as_double(0x7ff0000000000000LL));
    A.t_final[i] = t[u];
#pragma unroll
    for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + i] = p[u][n];
    const unsigned long long attempts = (unsigned long long)n_steps[u] + n_rej[u];
    A.steps[i] = n_steps[u];
    A.rejected[i] = n_rej[u];
    // as the reference evaluates: S per attempt, k[0] once at the start and once after every undo
    A.rhs_evals[i] = attempts > 0 ? attempts * S + 1 + n_rej[u] : 0;
    A.status[i] = flags[u] | (finite ? 0u : uint32_t(DFB_TRAJ_NONFINITE));
    if (TRACE) A.n_samples[i] = n_samp[u];
  }

  // one attempted step for every slot (make_step + StepsizeController::update + commit/undo)
  __device__ __forceinline__ void attempt() {
    double hs[TPT];
    double k[TPT][S][N];
    double y[TPT][N];
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      hs[u] = fmin(h_auto[u], A.t1 - t[u]);  // solver.rs:293
#pragma unroll
      for (int n = 0; n < N; ++n) k[u][0][n] = d[u][n];
    }
#pragma unroll
    for (int i = 1; i < S; ++i) {
#pragma unroll
      for (int u = 0; u < TPT; ++u) {
#pragma unroll
        for (int n = 0; n < N; ++n) {
          double acc = 0.0;
          bool first = true;
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (AMASK & amask_bit(i, j)) {
              acc = first ? k[u][j][n] * A.a[i * DFB_MAX_STAGES + j]
                          : fma(k[u][j][n], A.a[i * DFB_MAX_STAGES + j], acc);
              first = false;
            }
          y[u][n] = fma(acc, hs[u], p[u][n]);
        }
        rhs(u, fma(A.c[i], hs[u], t[u]), y[u], k[u][i]);
      }
    }
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      double ae[N], den[N];
      bool den_ok = true;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = 0.0, ea = 0.0;
        bool first = true, efirst = true;
#pragma unroll
        for (int j = 0; j < S; ++j) {
          if (BMASK & (1u << j)) {
            acc = first ? k[u][j][n] * A.b[j] : fma(k[u][j][n], A.b[j], acc);
            first = false;
          }
          if (EMASK & (1u << j)) {
            ea = efirst ? k[u][j][n] * A.e[j] : fma(k[u][j][n], A.e[j], ea);
            efirst = false;
          }
        }
        y[u][n] = fma(acc, hs[u], p[u][n]);
        // AutomaticStepsize::update (stepsize.rs:65-71): |e| / (atol + |e| rtol), max over components
        ae[n] = fabs(ea * hs[u]);
        den[n] = fma(ae[n], A.ctl.rtol[n], A.ctl.atol[n]);
        den_ok = den_ok && safely_normal(den[n]);
      }
      const double t_new = t[u] + hs[u];
      double d_new[N];
      rhs(u, t_new, y[u], d_new);

      double err, factor;
      bool rejected;
      bool fast = den_ok;
      if (fast) {
        // every quotient is finite and non-negative here, so max is a compare-select and
        // err >= 1 can be read off the exponent field
        err = ae[0] * fast_rcp_pos(den[0]);
#pragma unroll
        for (int n = 1; n < N; ++n) {
          const double q = ae[n] * fast_rcp_pos(den[n]);
          err = q > err ? q : err;
        }
        fast = in_root_range(err);
      }
      if (fast) {
        rejected = unsigned(__double2hiint(err)) >= 0x3ff00000u;
        factor = A.ctl.fac * inv_root_fast<P>(err, A.p, A.inv_p, A.inv_pf);
      } else {  // the reference's formula, IEEE division and pow (err = 0, inf, NaN, extreme tolerances)
        err = ae[0] / den[0];
#pragma unroll
        for (int n = 1; n < N; ++n) err = fmax(err, ae[n] / den[n]);
        rejected = err >= 1.0;
        factor = A.ctl.fac * pow(1.0 / err, A.inv_p);
      }
      factor = dev_clamp(factor, A.ctl.fac_min, A.ctl.fac_max);
      double h_next = h_auto[u] * factor;
      h_next = dev_clamp(h_next, A.ctl.stepsize_min, A.ctl.stepsize_max);
      // Rejections are rare (~0.1 % of the attempts of the Lorenz ensemble): when no lane of the
      // warp rejects, the commit is a plain register rename instead of 2N+2 selects per lane.
      // (A lane's own rejection is always part of the vote it takes part in.)
      if (__any_sync(__activemask(), rejected)) {
        if (rejected) {
          ++n_rej[u];
          // A rejection that leaves the controller's stepsize unchanged (clamped at
          // stepsize_range.start) repeats the same step from the same state with the same error:
          // the reference would spin forever here (solver.rs:294-297).  Stop the trajectory.
          if (h_next == h_auto[u]) {
            flags[u] |= DFB_TRAJ_STEPSIZE_FLOOR;
            t[u] = __longlong_as_double(0x7ff8000000000000LL);
          }
        } else {
          ++n_steps[u];
          t[u] = t_new;
#pragma unroll
          for (int n = 0; n < N; ++n) {
            p[u][n] = y[u][n];
            d[u][n] = d_new[n];
          }
          sample(u);
        }
      } else {
        ++n_steps[u];
        t[u] = t_new;
#pragma unroll
        for (int n = 0; n < N; ++n) {
          p[u][n] = y[u][n];
          d[u][n] = d_new[n];
        }
        sample(u);
      }
      h_auto[u] = h_next;
    }
  }

  __device__ void run() {
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      active[u] = false;
      t[u] = 0.0;
      h_auto[u] = 0.0;
#pragma unroll
      for (int n = 0; n < N; ++n) p[u][n] = d[u][n] = 0.0;
#pragma unroll
      for (int n = 0; n < NP; ++n) prm[u][n] = 0.0;
      fetch(u);
    }
    while (true) {
      bool any = false;
#pragma unroll
      for (int u = 0; u < TPT; ++u) {
        while (active[u] && !(t[u] < A.t1)) {  // solver.rs:292
          store(u);
          fetch(u);
        }
        any = any || active[u];
      }
      if (!any) break;
      attempt();
    }
  }
};

template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, unsigned EMASK, int BLOCK, int TPT, int P,
          bool TRACE = false>
__global__ void __launch_bounds__(BLOCK)
    ode_adaptive_kernel(const __grid_constant__ OdeAdaptiveArgs A) {
  OdeAdaptiveStepper<Sys, S, AMASK, BMASK, EMASK, TPT, P, TRACE> st(A);
  st.run();
}

}  // namespace DFB_NS
