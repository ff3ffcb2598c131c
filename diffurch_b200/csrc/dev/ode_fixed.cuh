// Hot path #1: fixed-step explicit RK over an ODE ensemble, no events, no
// delay history (reference: Solver::run solver.rs:292-310 with a scalar
// stepsize, State::make_step state.rs:94-120).
//
// One trajectory per thread.  p_prev, p_curr, d_curr and the S stage vectors
// k[S][N] live in registers for the whole integration; HBM is touched only to
// read y0/params and to write the final state, so the kernel is bound by the
// FP64 pipe.  The tableau comes in through the kernel parameter block (constant
// bank operands); its SPARSITY is a template argument so structurally-zero a_ij /
// b_j cost no instruction.  In FAST arithmetic the stage sums use rows
// pre-multiplied by h (one FMA per non-zero entry, accumulated onto p_prev);
// STRICT keeps the reference's order  p_prev + (sum_j k_j a_ij) * h  with every
// multiply and add rounded separately.
#pragma once
#include "common.cuh"
#include "systems.cuh"

#ifndef DFB_ODE_FOLD_MIN_N
#define DFB_ODE_FOLD_MIN_N 8
#endif

namespace DFB_NS {

// history stub for ODE right-hand sides (they never call p(t)/d(t))
struct NoHist {
  template <int D>
  __device__ __forceinline__ void eval(double, double*) {}
};

struct OdeFixedArgs {
  double a[DFB_MAX_STAGES * DFB_MAX_STAGES];   // a_ij
  double b[DFB_MAX_STAGES];
  double c[DFB_MAX_STAGES];
  double ha[DFB_MAX_STAGES * DFB_MAX_STAGES];  // h * a_ij (host-computed, FAST only)
  double hb[DFB_MAX_STAGES];                   // h * b_j
  double hc[DFB_MAX_STAGES];                   // c_i * h  (same rounding as the reference's c[i]*t_step)
  double t0, t1, h;
  size_t n_traj;
  const double* y0;      // [N][n]
  const double* params;  // [P][n]
  double* t_final;       // [n]
  double* y_final;       // [N][n]
  unsigned long long* steps;
  unsigned long long* rhs_evals;
  uint32_t* status;
  // ---- time-only events (ode_fixed_periodic_kernel): Periodic locators + their callbacks
  int32_t n_loc;
  int32_t event_capacity;
  dfb_locator loc[DFB_MAX_LOCATORS];
  double* aux_final;             // [A][n]
  unsigned long long* events;    // [n]
  double* ev_t;                  // [ecap][n]
  int32_t* ev_idx;               // [ecap][n]
  uint32_t* n_events;            // [n]
  // ---- state-function events (ode_event_kernel, dev/ode_event.cuh)
  double bi[DFB_MAX_STAGES * DFB_MAX_INTERP];  // continuous extension (rk.rs:13)
  long long max_steps;
  unsigned long long* queue;     // work counter (zeroed before the launch)
  unsigned long long* locate_iters;  // [n] location-loop iterations per member (may be null)
  int32_t locate_batch, locate_wait;
  int32_t inner_steps;           // step rounds between two warp votes (>= 1)
  int32_t on_start_cb;           // Solver::on_start_mut: callback id run on the initial State (0 = none)
  int32_t trace_every, trace_capacity;
  double* trace_t;               // [cap][n]
  double* trace_y;               // [cap][N][n]
  uint32_t* n_samples;           // [n]
};

__host__ __device__ constexpr unsigned long long amask_bit(int i, int j) { return 1ull << (i * 8 + j); }
__host__ __device__ constexpr unsigned long long amask_full_lower(int S) {
  unsigned long long m = 0;
  for (int i = 1; i < S; ++i)
    for (int j = 0; j < i; ++j) m |= amask_bit(i, j);
  return m;
}

// TPT = trajectories per thread: their stage arithmetic is interleaved, which
// doubles the independent DFMA chains per warp and amortises the non-FP64
// instructions (uniform-register traffic for the tableau, loop control, the
// shared time grid) over TPT trajectories.
template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, bool PRESCALED, int TPT>
struct OdeFixedStepper {
  static constexpr int N = Sys::N;
  static constexpr int NP = Sys::NP > 0 ? Sys::NP : 1;
  using Ref = StateRef<N, NoHist>;

  const OdeFixedArgs& A;
  double t_curr, t_prev;  // the time grid is shared by every trajectory (fixed step)
  double p_curr[TPT][N], p_prev[TPT][N], d_curr[TPT][N];
  double k[TPT][S][N];
  double prm[TPT][NP];
  NoHist nohist;

  __device__ explicit OdeFixedStepper(const OdeFixedArgs& A_) : A(A_) {}

  __device__ __forceinline__ void rhs_all(double (*out)[N]) {
#pragma unroll
    for (int u = 0; u < TPT; ++u)
      Sys::rhs(Ref{t_curr, p_curr[u], d_curr[u], t_prev, p_prev[u], &nohist}, prm[u], out[u]);
  }
  __device__ __forceinline__ void rhs_stage(int i) {
#pragma unroll
    for (int u = 0; u < TPT; ++u)
      Sys::rhs(Ref{t_curr, p_curr[u], d_curr[u], t_prev, p_prev[u], &nohist}, prm[u], k[u][i]);
  }

  // One make_step of length t_step; FULL = (t_step == A.h), the only case the
  // pre-scaled rows are valid for.
  // Wide states (N * TPT >= DFB_ODE_FOLD_MIN_N doubles per stage vector) do not keep all S stage
  // derivatives: after k[0 .. M-1], M = ceil(S/2), every later row of the tableau (and the b row) is
  // opened as a running sum over those, and each later derivative is folded into the rows that still
  // need it the moment the right-hand side returns it.  Same operations in the same order per
  // element as the plain form below (bit-identical, STRICT included) — only their position in time
  // changes: at most ~S/2 + 1 stage vectors are live instead of S + 1 (N = 12, rktp64: 218 -> ~150
  // registers).  Nothing after the step reads k[1 .. S-1] in the kernels that use this stepper.
  static constexpr bool FOLD = (N * TPT >= DFB_ODE_FOLD_MIN_N) && S >= 4;

  // FMA form of a folded step: PRESCALED builds (FAST) run every row as acc = fma(k, coef, acc) starting
  // from p_prev — full steps with the pre-scaled rows h a_ij / h b_j, any other step length (the step to
  // an event time, the last step of the interval) with the derivatives scaled by t_step as they are
  // produced and the plain rows a_ij / b_j.  p_prev is then dead once the rows are open.  STRICT builds
  // keep the reference's p_prev + (sum a_ij k_j) * t_step.
  template <bool FULL>
  __device__ __forceinline__ double row_coef(int r, int j) const {
    if (PRESCALED && FULL) return r < S ? A.ha[r * DFB_MAX_STAGES + j] : A.hb[j];
    return r < S ? A.a[r * DFB_MAX_STAGES + j] : A.b[j];
  }
  __device__ __forceinline__ static constexpr bool row_has(int r, int j) {
    return r < S ? (AMASK & amask_bit(r, j)) != 0ull : (BMASK & (1u << j)) != 0u;
  }
  template <bool FULL>
  __device__ __forceinline__ double row_open(int r, int u, int n, int m_stored) {
    // running sum of row r (r < S: stage row, r == S: the b row) over k[0 .. m_stored-1]
    double acc = PRESCALED ? p_prev[u][n] : 0.0;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      if (j >= m_stored) break;
      if (!row_has(r, j)) continue;
      if (PRESCALED) acc = fma(k[u][j][n], row_coef<FULL>(r, j), acc);
      else acc = acc + k[u][j][n] * row_coef<FULL>(r, j);
    }
    return acc;
  }

  template <bool FULL>
  __device__ __forceinline__ void step_fold(double t_step) {
    constexpr int M = (S + 1) / 2;
    constexpr bool SCALE = PRESCALED && !FULL;
    double row[TPT][S + 1][N];  // rows M+1 .. S only (the rest is never touched and takes no registers)
    double kj[TPT][N];
    t_prev = t_curr;
#pragma unroll
    for (int u = 0; u < TPT; ++u)
#pragma unroll
      for (int n = 0; n < N; ++n) {
        k[u][0][n] = SCALE ? d_curr[u][n] * t_step : d_curr[u][n];
        p_prev[u][n] = p_curr[u][n];
      }
#pragma unroll
    for (int i = 1; i < S; ++i) {
      t_curr = t_prev + (FULL ? A.hc[i] : A.c[i] * t_step);
#pragma unroll
      for (int u = 0; u < TPT; ++u)
#pragma unroll
        for (int n = 0; n < N; ++n) {
          const double acc = i <= M ? row_open<FULL>(i, u, n, i) : row[u][i][n];
          p_curr[u][n] = PRESCALED ? acc : p_prev[u][n] + acc * t_step;
          if (i == M) {
#pragma unroll
            for (int r = M + 1; r <= S; ++r) row[u][r][n] = row_open<FULL>(r, u, n, M);
          }
        }
      if (i < M) {
        rhs_stage(i);
        if (SCALE) {
#pragma unroll
          for (int u = 0; u < TPT; ++u)
#pragma unroll
            for (int n = 0; n < N; ++n) k[u][i][n] *= t_step;
        }
      } else {
        rhs_all(kj);
#pragma unroll
        for (int u = 0; u < TPT; ++u)
#pragma unroll
          for (int n = 0; n < N; ++n) {
            const double kv = SCALE ? kj[u][n] * t_step : kj[u][n];
#pragma unroll
            for (int r = i + 1; r <= S; ++r) {
              if (!row_has(r, i)) continue;
              if (PRESCALED) row[u][r][n] = fma(kv, row_coef<FULL>(r, i), row[u][r][n]);
              else row[u][r][n] = row[u][r][n] + kv * row_coef<FULL>(r, i);
            }
          }
      }
    }
#pragma unroll
    for (int u = 0; u < TPT; ++u)
#pragma unroll
      for (int n = 0; n < N; ++n)
        p_curr[u][n] = PRESCALED ? row[u][S][n] : p_prev[u][n] + row[u][S][n] * t_step;
    t_curr = t_prev + t_step;
    rhs_all(d_curr);
  }

  template <bool FULL>
  __device__ __forceinline__ void step(double t_step) {
    if constexpr (FOLD) {
      step_fold<FULL>(t_step);
      return;
    }
    // k[0] = d_curr (state.rs:95-96); the caller guarantees d_curr is current
    t_prev = t_curr;
#pragma unroll
    for (int u = 0; u < TPT; ++u)
#pragma unroll
      for (int n = 0; n < N; ++n) {
        k[u][0][n] = d_curr[u][n];
        p_prev[u][n] = p_curr[u][n];
      }
#pragma unroll
    for (int i = 1; i < S; ++i) {
      t_curr = t_prev + (FULL ? A.hc[i] : A.c[i] * t_step);
#pragma unroll
      for (int u = 0; u < TPT; ++u)
#pragma unroll
        for (int n = 0; n < N; ++n) {
          if (PRESCALED && FULL) {
            double acc = p_prev[u][n];
#pragma unroll
            for (int j = 0; j < i; ++j)
              if (AMASK & amask_bit(i, j)) acc = fma(k[u][j][n], A.ha[i * DFB_MAX_STAGES + j], acc);
            p_curr[u][n] = acc;
          } else {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < i; ++j)
              if (AMASK & amask_bit(i, j)) acc = acc + k[u][j][n] * A.a[i * DFB_MAX_STAGES + j];
            p_curr[u][n] = p_prev[u][n] + acc * t_step;
          }
        }
      rhs_stage(i);
    }
#pragma unroll
    for (int u = 0; u < TPT; ++u)
#pragma unroll
      for (int n = 0; n < N; ++n) {
        if (PRESCALED && FULL) {
          double acc = p_prev[u][n];
#pragma unroll
          for (int j = 0; j < S; ++j)
            if (BMASK & (1u << j)) acc = fma(k[u][j][n], A.hb[j], acc);
          p_curr[u][n] = acc;
        } else {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < S; ++j)
            if (BMASK & (1u << j)) acc = acc + k[u][j][n] * A.b[j];
          p_curr[u][n] = p_prev[u][n] + acc * t_step;
        }
      }
    t_curr = t_prev + t_step;
    rhs_all(d_curr);
  }

  // gid[u] = trajectory of slot u (clamped for the ragged tail; `valid` masks the stores)
  // Output policy standing in for an `on_step` closure (solver.rs:290, 309): every k-th invocation
  // stores (t, p) into the [capacity][1 + N][n] trace, one coalesced row per component.  The time
  // grid is shared, so whether this invocation is sampled is the same for the whole grid.
  __device__ __forceinline__ void on_step(const size_t* gid, const bool* valid, unsigned long long calls,
                                          uint32_t& n_samples) {
    if (calls % (unsigned long long)A.trace_every == 0 && n_samples < uint32_t(A.trace_capacity)) {
      const size_t n_traj = A.n_traj;
#pragma unroll
      for (int u = 0; u < TPT; ++u) {
        if (!valid[u]) continue;
        A.trace_t[size_t(n_samples) * n_traj + gid[u]] = t_curr;
#pragma unroll
        for (int n = 0; n < N; ++n) A.trace_y[(size_t(n_samples) * N + n) * n_traj + gid[u]] = p_curr[u][n];
      }
      ++n_samples;
    }
  }

  template <bool TRACE = false>
  __device__ void run(const size_t* gid, const bool* valid) {
    const size_t n_traj = A.n_traj;
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
#pragma unroll
      for (int n = 0; n < Sys::NP; ++n) prm[u][n] = A.params[size_t(n) * n_traj + gid[u]];
#pragma unroll
      for (int n = 0; n < N; ++n) {
        p_curr[u][n] = p_prev[u][n] = A.y0[size_t(n) * n_traj + gid[u]];
        d_curr[u][n] = 0.0;
      }
    }
    t_curr = t_prev = A.t0;
    unsigned long long n_steps = 0;
    uint32_t n_samples = 0u;
    if (TRACE) on_step(gid, valid, 0ull, n_samples);  // events_on_step at t_init (solver.rs:290)
    if (t_curr < A.t1) {
      rhs_all(d_curr);  // first make_step evaluates k[0] itself (state.rs:97-98)
      const double h = A.h;
      // Solver::run requests min(h, t_end - t_curr) (solver.rs:293): that is a full
      // step exactly while t_end - t_curr >= h.  Full steps and the closing partial
      // step(s) get separate loops so each keeps only its own constants live.
#pragma unroll 1
      while (t_curr < A.t1 && (A.t1 - t_curr) >= h) {
        step<true>(h);
        ++n_steps;
        if (TRACE) on_step(gid, valid, n_steps, n_samples);
      }
#pragma unroll 1
      while (t_curr < A.t1) {
        step<false>(fmin(h, A.t1 - t_curr));
        ++n_steps;
        if (TRACE) on_step(gid, valid, n_steps, n_samples);
      }
    }
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      if (!valid[u]) continue;
      bool finite = true;
#pragma unroll
      for (int n = 0; n < N; ++n)
        finite = finite && (fabs(p_curr[u][n]) < __longlong_as_double(0x7ff0000000000000LL));
      A.t_final[gid[u]] = t_curr;
#pragma unroll
      for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + gid[u]] = p_curr[u][n];
      A.steps[gid[u]] = n_steps;
      A.rhs_evals[gid[u]] = n_steps > 0 ? n_steps * S + 1 : 0;
      A.status[gid[u]] = finite ? 0u : uint32_t(DFB_TRAJ_NONFINITE);
      if (TRACE && A.n_samples) A.n_samples[gid[u]] = n_samples;
    }
  }
};

template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, bool PRESCALED, int BLOCK, int TPT,
          bool TRACE = false>
__global__ void __launch_bounds__(BLOCK)
    ode_fixed_kernel(const __grid_constant__ OdeFixedArgs A) {
  // slot u of thread g handles trajectory g + u * (threads in the grid): coalesced for every u.
  // BLOCK is the launch bound; the host sizes the block (<= BLOCK) from the ensemble (host/launch_geometry.hpp).
  const size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t gid[TPT];
  bool valid[TPT];
#pragma unroll
  for (int u = 0; u < TPT; ++u) {
    const size_t i = g + size_t(u) * stride;
    valid[u] = i < A.n_traj;
    gid[u] = valid[u] ? i : A.n_traj - 1;
  }
  if (!valid[0]) return;
  OdeFixedStepper<Sys, S, AMASK, BMASK, PRESCALED, TPT> st(A);
  st.template run<TRACE>(gid, valid);
}

// ---------------------------------------------------------------------------------------
// Fixed step + time-only events: every locator is a `Periodic` (loc.rs:292-335), the shape of
// every Lyapunov-exponent run in the reference (tests/lyapunov_exponents.rs: integrate the
// variational system, re-orthonormalise with a QR every `period`).  Detection and location of
// a Periodic depend on the time grid only, and with a fixed step the time grid is the same for
// every member of the ensemble, so the whole event machinery of Solver::run (solver.rs:299-307:
// locate earliest -> undo -> step to the event time -> commit -> zero step -> callback ->
// commit) is WARP-UNIFORM control flow around the register-resident stepper above: no lane
// parks, no divergence, and the state never leaves registers.  Only the callback's
// `stop_integration` is per lane (a stopped lane stores its result and idles).
template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, bool PRESCALED, int TPT>
struct OdeFixedPeriodicStepper : OdeFixedStepper<Sys, S, AMASK, BMASK, PRESCALED, TPT> {
  using Base = OdeFixedStepper<Sys, S, AMASK, BMASK, PRESCALED, TPT>;
  static constexpr int N = Sys::N;
  static constexpr int NA = Sys::NA > 0 ? Sys::NA : 1;
  using RefMut = StateRefMut<N, NoHist>;
  using Base::A;
  using Base::d_curr;
  using Base::k;
  using Base::p_curr;
  using Base::p_prev;
  using Base::prm;
  using Base::t_curr;
  using Base::t_prev;

  double aux[TPT][NA];
  bool alive[TPT];

  __device__ explicit OdeFixedPeriodicStepper(const OdeFixedArgs& A_) : Base(A_) {}

  __device__ void store(int u, size_t gid, double t_fin, unsigned long long n_steps,
                        unsigned long long n_events, unsigned long long n_rhs, uint32_t n_evlog) {
    const size_t n_traj = A.n_traj;
    bool finite = true;
#pragma unroll
    for (int n = 0; n < N; ++n)
      finite = finite && (fabs(p_curr[u][n]) < __longlong_as_double(0x7ff0000000000000LL));
    uint32_t status = finite ? 0u : uint32_t(DFB_TRAJ_NONFINITE);
    if (t_fin == __longlong_as_double(0x7ff0000000000000LL)) status |= DFB_TRAJ_STOPPED;
    A.t_final[gid] = t_fin;
#pragma unroll
    for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + gid] = p_curr[u][n];
    if (A.aux_final) {
#pragma unroll
      for (int n = 0; n < Sys::NA; ++n) A.aux_final[size_t(n) * n_traj + gid] = aux[u][n];
    }
    A.steps[gid] = n_steps;
    A.events[gid] = n_events;
    A.rhs_evals[gid] = n_rhs;
    A.status[gid] = status;
    if (A.n_events) A.n_events[gid] = n_evlog;
  }

  __device__ void run(const size_t* gid, const bool* valid) {
    const size_t n_traj = A.n_traj;
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
#pragma unroll
      for (int n = 0; n < Sys::NP; ++n) prm[u][n] = A.params[size_t(n) * n_traj + gid[u]];
#pragma unroll
      for (int n = 0; n < N; ++n) {
        p_curr[u][n] = p_prev[u][n] = A.y0[size_t(n) * n_traj + gid[u]];
        d_curr[u][n] = 0.0;
      }
#pragma unroll
      for (int n = 0; n < NA; ++n) aux[u][n] = 0.0;
      alive[u] = valid[u];
    }
    t_curr = t_prev = A.t0;
    // uniform bookkeeping (identical in every lane)
    unsigned long long n_steps = 0, n_events = 0, n_rhs = 0;
    uint32_t n_evlog = 0u;
    double last_call[DFB_MAX_LOCATORS];
    uint32_t has_last_call = 0u;
#pragma unroll
    for (int l = 0; l < DFB_MAX_LOCATORS; ++l) last_call[l] = 0.0;
    const double h = A.h;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double t_quiet = -inf;

#pragma unroll 1
    while (t_curr < A.t1) {
      bool any_alive = false;
#pragma unroll
      for (int u = 0; u < TPT; ++u) any_alive = any_alive || alive[u];
      if (!__any_sync(0xffffffffu, any_alive)) break;  // every member of the warp stopped itself

      if (t_prev == t_curr) {  // state.rs:97-98: k[0] evaluated by make_step itself
        this->rhs_all(d_curr);
        ++n_rhs;
      }
      const double h_req = fmin(h, A.t1 - t_curr);  // solver.rs:293
      const double t_next = t_curr + h_req;         // = t_curr after make_step(h_req) (state.rs:117)

      // HListLocateEarliest over Periodic locators (loc.rs:318-334, 821-835, 1002-1011).  Detection
      // and location read the time grid only, so they run BEFORE the step on (t_curr, t_next): when an
      // event is found the reference's full step would be undone (solver.rs:300) and is not taken at
      // all — its S + 1 right-hand-side calls are counted, nothing else of it survives (d_curr is
      // re-evaluated from the same state to the same bits).  `t_quiet` is a lower bound of the next
      // period boundary of any locator, a wide margin (1e-9 relative, ~10^7 ulp) below it: while t_next
      // stays under it floor((t - o) / T) cannot have moved, and the two FP64 divisions per locator
      // and step are skipped; the exact test decides inside the margin.
      bool found = false;
      double best_t = 0.0;
      int best_l = 0;
      if (!(t_next < t_quiet)) {
        double quiet = inf;
#pragma unroll
        for (int l = 0; l < DFB_MAX_LOCATORS; ++l) {
          if (l >= A.n_loc) break;
          const dfb_locator& L = A.loc[l];
          const double fl_prev = floor((t_curr - L.offset) / L.period);
          const double fl_curr = floor((t_next - L.offset) / L.period);
          const double bound = (fl_curr + 1.0) * L.period + L.offset;
          const double q = bound - 1e-9 * (fabs(bound) + fabs(L.offset) + fabs(L.period));
          // (a locator without DedupLocF may fire again on the step that starts AT its event time when
          // floor() rounds down there — solver.rs does — so only deduplicated locators may be skipped)
          const bool skippable = L.period > 0.0 && L.dedup != 0;
          quiet = skippable ? (q < quiet ? q : quiet) : -inf;
          if (L.dedup && ((has_last_call >> l) & 1u) && !(t_curr > last_call[l])) continue;
          if (fl_prev < fl_curr) {
            const double tl = fl_curr * L.period + L.offset;
            if (!found || tl < best_t) {
              found = true;
              best_t = tl;
              best_l = l;
            }
          }
        }
        // the bounds assume the whole step (t_curr, t_next] is taken; when an event cuts it short another
        // locator's boundary may still lie ahead inside it: decide exactly again after the event
        t_quiet = found ? -inf : quiet;
      }
      if (found && best_t >= t_curr) {  // solver.rs:299-307
        n_rhs += S + 1;  // the undone full step and the k[0] make_step re-evaluates after undo_step
        if (!Sys::pure_rhs) {  // that re-evaluation sees t_prev == t_curr, p_prev == p_curr (state.rs:146-150)
          t_prev = t_curr;
#pragma unroll
          for (int u = 0; u < TPT; ++u)
#pragma unroll
            for (int n = 0; n < N; ++n) p_prev[u][n] = p_curr[u][n];
          this->rhs_all(d_curr);
        }
        this->template step<false>(best_t - t_curr);
        n_rhs += S;
        // commit_step; make_zero_step (state.rs:141-144)
        t_prev = t_curr;
#pragma unroll
        for (int u = 0; u < TPT; ++u)
#pragma unroll
          for (int n = 0; n < N; ++n) p_prev[u][n] = p_curr[u][n];
        // callback of the located event (LocCallback / DedupLocF::eval_mut, loc.rs:1023-1026)
        const dfb_locator& L = A.loc[best_l];
        if (L.dedup) {
          has_last_call |= 1u << best_l;
          last_call[best_l] = t_curr;
        }
        ++n_events;
        if (A.event_capacity > 0) {
          if (n_evlog < uint32_t(A.event_capacity)) {
#pragma unroll
            for (int u = 0; u < TPT; ++u)
              if (alive[u]) {
                A.ev_t[size_t(n_evlog) * n_traj + gid[u]] = best_t;
                A.ev_idx[size_t(n_evlog) * n_traj + gid[u]] = best_l;
              }
          }
          ++n_evlog;
        }
        ++n_steps;  // the second commit_step of the event (solver.rs:308)
        if (L.callback_id != DFB_CB_NONE) {
#pragma unroll
          for (int u = 0; u < TPT; ++u) {
            double t_lane = t_curr;
            RefMut m{&t_lane, p_curr[u], d_curr[u], t_prev, p_prev[u], &this->nohist};
            Sys::callback(L.callback_id, m, prm[u], aux[u]);
            if (t_lane != t_curr && alive[u]) {  // stop_integration (state_fn.rs:92-97) or a moved clock
              store(u, gid[u], t_lane, n_steps, n_events, n_rhs, n_evlog);
              alive[u] = false;
            }
          }
        }
      } else {
        if (h_req == h) this->template step<true>(h);
        else this->template step<false>(h_req);
        n_rhs += S;
        ++n_steps;
      }
    }
#pragma unroll
    for (int u = 0; u < TPT; ++u)
      if (alive[u]) store(u, gid[u], t_curr, n_steps, n_events, n_rhs, n_evlog);
  }
};

template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, bool PRESCALED, int BLOCK, int TPT, int MINB = 1>
__global__ void __launch_bounds__(BLOCK, MINB)
    ode_fixed_periodic_kernel(const __grid_constant__ OdeFixedArgs A) {
  // BLOCK is the launch bound; the host sizes the block (<= BLOCK) from the ensemble (host/launch_geometry.hpp)
  const size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t gid[TPT];
  bool valid[TPT];
#pragma unroll
  for (int u = 0; u < TPT; ++u) {
    const size_t i = g + size_t(u) * stride;
    valid[u] = i < A.n_traj;
    gid[u] = valid[u] ? i : A.n_traj - 1;
  }
  // whole warps stay (the loop votes); lanes past the end replay the last member and store nothing
  if ((g & ~size_t(31)) >= A.n_traj) return;
  OdeFixedPeriodicStepper<Sys, S, AMASK, BMASK, PRESCALED, TPT> st(A);
  st.run(gid, valid);
}

}  // namespace DFB_NS
