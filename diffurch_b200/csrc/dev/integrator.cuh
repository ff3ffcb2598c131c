// General per-trajectory integrator: the device form of the reference's
// `Solver::run` (src/solver.rs:263-314) with every builder feature selectable at
// run time from the DevProblem block: fixed or automatic stepsize, up to MAXL
// locators of any detection/location kind, Periodic, discontinuity propagation,
// delay history.  One trajectory per thread; State (state.rs:26-42) in
// registers; the history ring in HBM as structure-of-arrays.
//
// The loop is written as a state machine with ONE make_step call site, so the
// lanes of a warp that re-take a step (rejected step, step truncated to an
// event time) run their stages together with the lanes that take an ordinary
// step.  Event location (bisection) is deferred: a lane that detects an event
// parks until enough lanes of its warp are parked (or nobody else can advance),
// then the parked lanes run their bisections together (warp-level batching).
#pragma once
#include "common.cuh"
#include "history_fn.cuh"
#include "systems.cuh"

namespace DFB_NS {

__device__ __forceinline__ double dev_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double dev_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// f64::powi (compiler-rt __powidf2 / LLVM ExpandPowI): square-and-multiply.
__device__ __forceinline__ double dev_powi(double a, int b) {
  double r = 1.0;
  const bool recip = b < 0;
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? 1.0 / r : r;
}

// n-th root in plain IEEE +, *, / (frexp / ldexp are exact): STRICT builds use it where the reference
// calls f64::powf (stepsize.rs:74) so that the adaptive path is bit-identical to the CPU oracle run with
// the same function (oracle/diffurch_oracle.hpp `portable_root`; libm pow differs between platforms in
// its last bit and CUDA's pow is not glibc's).  Within 2 ulp of pow(x, 1/n) for 1e-4 < x < 1e4.
__device__ inline double dev_portable_root(double x, unsigned n) {
  if (x != x) return x;
  if (x < 0.0) return __longlong_as_double(0x7ff8000000000000LL);
  if (x == 0.0) return 0.0;
  if (x == __longlong_as_double(0x7ff0000000000000LL)) return x;
  if (n <= 1u) return x;
  int e = 0;
  const double m = frexp(x, &e);
  int q = e / int(n), r = e - q * int(n);
  if (r < 0) {
    r += int(n);
    q -= 1;
  }
  const double a = ldexp(m, r);
  double y = a > 1.0 ? a : 1.0;
  const double nm1 = double(n - 1u), dn = double(n);
#pragma unroll 1
  for (int it = 0; it < 48; ++it) {
    double p = 1.0;
#pragma unroll 1
    for (unsigned j = 0; j + 1u < n; ++j) p = p * y;
    y = (nm1 * y + a / p) / dn;
  }
  return ldexp(y, q);
}

// Tableau accessor over the kernel parameter block, compile-time shape.
// RT = run-time shape: the kernel is instantiated for the maximal shape (S_, I_ = array sizes and
// unroll bounds) and every stage / interpolant loop iteration is additionally guarded by the
// tableau's own S and I.  One such instantiation per system is the catch-all that makes every
// (system, tableau) pair runnable; the exact-shape instantiations (RT = false) are the fast ones.
template <int S_, int I_, bool RT_ = false>
struct ParamTab {
  static constexpr int S = S_, I = I_;
  static constexpr bool RT = RT_;
  const dfb_tableau& t;
  __device__ __forceinline__ bool stage(int i) const { return !RT || i < t.S; }
  __device__ __forceinline__ bool term(int j) const { return !RT || j < t.I; }
  __device__ __forceinline__ int s_rt() const { return RT ? t.S : S; }
  __device__ __forceinline__ double a(int i, int j) const { return t.a[i * DFB_MAX_STAGES + j]; }
  __device__ __forceinline__ double b(int j) const { return t.b[j]; }
  __device__ __forceinline__ double b2(int j) const { return t.b2[j]; }
  __device__ __forceinline__ double c(int j) const { return t.c[j]; }
  __device__ __forceinline__ double bi(int i, int j) const { return t.bi[i * DFB_MAX_INTERP + j]; }
};

// ButcherTableu::dense_output::<D> (rk.rs:17-52), reference operation order.
template <int D, int N, class Tab>
__device__ __forceinline__ void dense_output(const Tab& tab, const double* y_prev, double t_step,
                                             double theta, const double (*k)[N], double* out) {
  constexpr int S = Tab::S, I = Tab::I;
  double pw[I];  // theta^j by powi's multiply sequence
  pw[0] = 1.0;
  if (I > 1) pw[1] = theta;
  if (I > 2) pw[2] = theta * theta;
  if (I > 3) pw[3] = theta * pw[2];
  if (I > 4) pw[4] = pw[2] * pw[2];
  double delta[N];
#pragma unroll
  for (int n = 0; n < N; ++n) delta[n] = 0.0;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    if (!tab.stage(i)) continue;
    double b_i = 0.0;
    if (D == 0) {
#pragma unroll
      for (int j = 0; j < I; ++j)
        if (tab.term(j)) b_i += tab.bi(i, j) * pw[j];
    } else {
#pragma unroll
      for (int j = 1; j < I; ++j)
        if (tab.term(j)) b_i += tab.bi(i, j) * pw[j - 1] * double(j);
    }
#pragma unroll
    for (int n = 0; n < N; ++n) delta[n] += k[i][n] * b_i;
  }
  if (D == 0) {
#pragma unroll
    for (int n = 0; n < N; ++n) out[n] = y_prev[n] + delta[n] * t_step;
  } else {
#pragma unroll
    for (int n = 0; n < N; ++n) out[n] = delta[n];
  }
}

// Partition point of a ring's time column (first record with t_r > t), started from `r`:
// forward while t_r <= t, backward while t_{r-1} > t, at most 8 probes, then bisection of what
// is left.  One out-of-line copy serves every lookup site of a kernel (the lookups of the stages,
// the detectors and the bisection iterations would otherwise each inline it).
static __device__ __noinline__ int history_partition_point(const double* t_col, size_t n_traj, int head, int mask,
                                                    int len, int r, double t) {
  auto T = [&](int q) { return t_col[size_t((head + q) & mask) * n_traj]; };
  int lo = 0, hi = len;
  int probes = 0;
  bool settled = false;
  while (probes < 8) {
    if (r < len && T(r) <= t) {
      lo = ++r;
      ++probes;
    } else {
      hi = r;
      break;
    }
  }
  if (probes < 8) {
    while (probes < 8) {
      if (r > lo && !(T(r - 1) <= t)) {
        hi = --r;
        ++probes;
      } else {
        settled = true;
        break;
      }
    }
  }
  if (!settled) {  // the start was far off (first lookup, state-dependent delay): bisect [lo, hi)
    while (lo < hi) {
      const int mid = lo + (hi - lo) / 2;
      if (T(mid) <= t) lo = mid + 1;
      else hi = mid;
    }
    r = lo;
  }
  return r;
}

// StateHistory (state.rs:11-23, 161-187) over a per-trajectory ring in HBM.
// Ring record r holds (t_r, y_r) and the stages k of the step that ENDED at
// t_r, i.e. k_deque[r-1] of the reference.
//
// Two things differ from the reference's VecDeque without changing what a lookup returns:
//  * the partition point (first record with t_r > t, state.rs:165) is found from a per-lane
//    HINT — the record the previous lookup landed in — by a short linear probe, falling back to
//    the binary search only when the probe runs long.  Delayed arguments move monotonically, so a
//    lookup costs ~2 loads of the time column instead of log2(len) dependent ones.  Same index as
//    the binary search (the column is sorted), so STRICT results are unchanged bit for bit;
//  * FAST arithmetic with `P->coeff_records`: the k slots of a record hold the pre-combined
//    coefficients of the step's interpolant, d_j = sum_i k_i bi[i][j] / h^(j-1) (j = 1..I-1), as in
//    dev/dde_fixed.cuh: a lookup loads N I doubles instead of N (1 + S) and is a Horner chain.
template <int N, class Tab, bool ENABLED>
struct History {
  static constexpr int S = Tab::S, I = Tab::I;
  const DevProblem* P;
  Tab tab;
  size_t gid;
  int head, len;
  uint32_t status;
  double y0[N];
  double hist_k;  // parameter of the sin history functions
  int base = 0;   // records pruned so far: absolute index of ring record 0
  int hint = 0;   // absolute index of the partition point of the previous lookup

  __device__ __forceinline__ size_t slot(int r) const {
    return size_t((head + r) & (P->ring_capacity - 1));
  }
  __device__ __forceinline__ double T(int r) const { return P->ring_t[slot(r) * P->n_traj + gid]; }

  // InitialCondition::eval::<D> (initial_condition.rs:15-50)
  template <int D>
  __device__ void init_eval(double t, double* out) {
    init_condition_eval<D, N>(P->history_id, t, y0, hist_k, status, out);
  }

  // VecDeque::partition_point(|t_i| t_i <= t): first record with t_r > t
  __device__ __forceinline__ int partition_point(double t) {
    int r = hint - base;
    r = r < 0 ? 0 : (r > len ? len : r);
    r = history_partition_point(P->ring_t + gid, P->n_traj, head, P->ring_capacity - 1, len, r, t);
    hint = base + r;
    return r;
  }

  template <int D>
  __device__ void eval(double t, double* out) {
    if (t <= P->t0) {
      init_eval<D>(t, out);
      return;
    }
    if (!ENABLED) {  // ODE systems never look anything up
#pragma unroll
      for (int n = 0; n < N; ++n) out[n] = dev_nan();
      status |= DFB_TRAJ_HISTORY_NOT_READY;
      return;
    }
    const int lo = partition_point(t);
    if (lo == 0 || lo == len) {
      status |= (lo == 0) ? DFB_TRAJ_HISTORY_DELETED : DFB_TRAJ_HISTORY_NOT_READY;
#pragma unroll
      for (int n = 0; n < N; ++n) out[n] = dev_nan();
      return;
    }
    const size_t n_traj = P->n_traj;
    const size_t s0 = slot(lo - 1), s1 = slot(lo);
    const double t_prev = P->ring_t[s0 * n_traj + gid];
#if DFB_ARITH == 1
    if (P->coeff_records) {
      const double s = t - t_prev;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const size_t Sr = size_t(tab.s_rt());
        double d[I > 1 ? I - 1 : 1];
#pragma unroll
        for (int j = 0; j < I - 1; ++j)
          d[j] = tab.term(j + 1) ? P->ring_k[((s1 * Sr + j) * N + n) * n_traj + gid] : 0.0;
        double acc = 0.0;
        bool started = false;
#pragma unroll
        for (int j = I - 2; j >= 0; --j) {
          if (!tab.term(j + 1)) continue;
          const double c = D == 0 ? d[j] : double(j + 1) * d[j];
          acc = started ? fma(acc, s, c) : c;
          started = true;
        }
        out[n] = D == 0 ? fma(acc, s, P->ring_y[(s0 * N + n) * n_traj + gid]) : acc;
      }
      return;
    }
#endif
    const double t_next = P->ring_t[s1 * n_traj + gid];
    double y_prev[N];
    double k[S][N];
#pragma unroll
    for (int n = 0; n < N; ++n) y_prev[n] = P->ring_y[(s0 * N + n) * n_traj + gid];
#pragma unroll
    for (int i = 0; i < S; ++i)
#pragma unroll
      for (int n = 0; n < N; ++n)
        k[i][n] = tab.stage(i) ? P->ring_k[((s1 * size_t(tab.s_rt()) + i) * N + n) * n_traj + gid] : 0.0;
    const double t_step = t_next - t_prev;
    const double theta = (t - t_prev) / t_step;
    dense_output<D, N>(tab, y_prev, t_step, theta, k, out);
  }

  // push of commit_step (state.rs:123-125); t_step = length of the step that ended at t
  __device__ void push(double t, const double* y, const double (*k)[N], double t_step) {
    if (!ENABLED) return;
    if (len == P->ring_capacity) {  // library limit, not a reference behaviour
      head = (head + 1) & (P->ring_capacity - 1);
      --len;
      ++base;
      status |= DFB_TRAJ_RING_OVERFLOW;
    }
    const size_t n_traj = P->n_traj;
    const size_t s = slot(len);
    P->ring_t[s * n_traj + gid] = t;
#pragma unroll
    for (int n = 0; n < N; ++n) P->ring_y[(s * N + n) * n_traj + gid] = y[n];
    if (k) {
#if DFB_ARITH == 1
      if (P->coeff_records) {
        const double inv = 1.0 / t_step;  // zero-length commits (solver.rs:308) give NaN records nobody reads
        double scale = 1.0;               // 1 / h^(j-1)
#pragma unroll
        for (int j = 1; j < I; ++j) {
          if (!tab.term(j)) continue;
#pragma unroll
          for (int n = 0; n < N; ++n) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < S; ++i)
              if (tab.stage(i)) acc = fma(k[i][n], tab.bi(i, j), acc);
            P->ring_k[((s * size_t(tab.s_rt()) + (j - 1)) * N + n) * n_traj + gid] = acc * scale;
          }
          scale *= inv;
        }
        ++len;
        return;
      }
#endif
#pragma unroll
      for (int i = 0; i < S; ++i)
#pragma unroll
        for (int n = 0; n < N; ++n)
          if (tab.stage(i)) P->ring_k[((s * size_t(tab.s_rt()) + i) * N + n) * n_traj + gid] = k[i][n];
    }
    ++len;
  }
  // prune of commit_step (state.rs:126-133)
  __device__ void prune(double t_tail) {
    if (!ENABLED) return;
    while (len > 1 && T(1) < t_tail) {
      head = (head + 1) & (P->ring_capacity - 1);
      --len;
      ++base;
    }
  }
};

// Per-trajectory discontinuity list (state.rs:20, 134-138) in HBM.
struct DiscoList {
  const DevProblem* P;
  size_t gid;
  int head, len;
  __device__ __forceinline__ size_t slot(int r) const {
    return size_t((head + r) & (DFB_DISCO_CAP - 1)) * P->n_traj + gid;
  }
  __device__ double t(int r) const { return P->disco_buf_t[slot(r)]; }
  __device__ int order(int r) const { return P->disco_buf_ord[slot(r)]; }
  __device__ bool push(double tt, int ord) {
    if (len == DFB_DISCO_CAP) return false;
    const size_t s = slot(len);
    P->disco_buf_t[s] = tt;
    P->disco_buf_ord[s] = ord;
    ++len;
    return true;
  }
  __device__ void prune(double t_tail) {
    while (len > 0 && t(0) < t_tail) {
      head = (head + 1) & (DFB_DISCO_CAP - 1);
      --len;
    }
  }
  // util::partition_point_linear (src/util/partition_point.rs:26-52), pred = t_i <= x
  __device__ int partition_point_linear(int start, double x) const {
    if (len == 0) return 0;
    int i = start < len - 1 ? start : len - 1;
    if (t(i) <= x) {
      i += 1;
      while (i < len && t(i) <= x) i += 1;
    } else {
      while (i > 0 && !(t(i - 1) <= x)) i -= 1;
    }
    return i;
  }
};

enum : int { MODE_STEP = 0, MODE_EVENT_RESTEP = 1, MODE_LOCATE = 2, MODE_DONE = 3 };
// status bits that stand for a `panic!` of the reference: the trajectory ends where it was raised
constexpr uint32_t kPanicBits = DFB_TRAJ_HISTORY_DELETED | DFB_TRAJ_HISTORY_NOT_READY | DFB_TRAJ_NO_DERIVATIVE;

// ODEHIST: keep the delay-history ring for a system that never looks anything up itself.  The
// reference stores the last two steps even for ODEs (commit_step with max_delay = 0 keeps three
// records, state.rs:122-133), and `State::eval` falls back to them for a time outside the step in
// flight (state.rs:83-92).  Only the RegulaFalsi location method can ask for such a time (its secant
// may leave the bracket, loc.rs:260-285), so only problems that use it pay for the ring.
template <class Sys, int S, int I, bool HAS_LOC, bool RT = false, bool ODEHIST = false>
struct Integrator {
  static constexpr int N = Sys::N;
  static constexpr int NP = Sys::NP > 0 ? Sys::NP : 1;
  static constexpr int NA = Sys::NA > 0 ? Sys::NA : 1;
  static constexpr int MAXL = HAS_LOC ? DFB_MAX_LOCATORS : 1;
  using Tab = ParamTab<S, I, RT>;
  using Hist = History<N, Tab, Sys::uses_history || ODEHIST>;
  using Ref = StateRef<N, Hist>;
  using RefMut = StateRefMut<N, Hist>;

  const DevProblem& P;
  Tab tab;
  size_t gid;
  // State (state.rs:26-42)
  double t_curr, t_prev;
  double p_curr[N], p_prev[N], d_curr[N], d_prev[N], e_curr[N];
  double k[S][N];
  double prm[NP];
  double aux[NA];
  Hist hist;
  DiscoList disco;
  // stepsize controller state
  double h_auto;
  // locator state (indexed at run time: lives in local memory, L1-resident)
  double last_call[MAXL];
  uint32_t has_last_call;
  int every_i[MAXL];
  double sep_prev[MAXL];
  double trk_t[MAXL];
  long long trk_order[MAXL];
  int trk_index[MAXL];
  // counters
  unsigned long long n_steps, n_rejected, n_events, n_rhs;
  uint32_t n_samples, n_evlog;
  unsigned long long on_step_calls;

  __device__ Integrator(const DevProblem& P_, size_t gid_)
      : P(P_), tab{P_.rk}, gid(gid_), hist{&P_, Tab{P_.rk}, gid_, 0, 0, 0u, {}, 0.0},
        disco{&P_, gid_, 0, 0} {}

  // ---- views (state_fn.rs:170-201) -----------------------------------------
  __device__ __forceinline__ Ref view_curr() { return Ref{t_curr, p_curr, d_curr, t_prev, p_prev, &hist}; }
  __device__ __forceinline__ Ref view_prev() { return Ref{t_prev, p_prev, d_prev, t_prev, p_prev, &hist}; }

  __device__ __forceinline__ void rhs_curr(double* out) {
    Sys::rhs(view_curr(), prm, out);
    ++n_rhs;
  }

  // State::eval::<D> (state.rs:83-92)
  template <int D>
  __device__ void state_eval(double t, double* out) {
    if (t >= t_prev && t <= t_curr) {
      const double t_step = t_curr - t_prev;
      const double theta = (t - t_prev) / t_step;
      dense_output<D, N>(tab, p_prev, t_step, theta, k, out);
    } else {
      hist.template eval<D>(t, out);
    }
  }

  // ---- State::make_step (state.rs:94-120) ----------------------------------
  __device__ __forceinline__ void make_step(double t_step, bool need_error) {
    if (t_prev != t_curr) {
#pragma unroll
      for (int n = 0; n < N; ++n) k[0][n] = d_curr[n];
    } else {
      rhs_curr(k[0]);
    }
    t_prev = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      p_prev[n] = p_curr[n];
      d_prev[n] = d_curr[n];
    }
#pragma unroll
    for (int i = 1; i < S; ++i) {
      if (!tab.stage(i)) continue;
      t_curr = t_prev + tab.c(i) * t_step;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < i; ++j) acc = acc + k[j][n] * tab.a(i, j);
        p_curr[n] = p_prev[n] + acc * t_step;
      }
      rhs_curr(k[i]);
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < S; ++j)
        if (tab.stage(j)) acc = acc + k[j][n] * tab.b(j);
      p_curr[n] = p_prev[n] + acc * t_step;
    }
    t_curr = t_prev + t_step;
    rhs_curr(d_curr);
    if (need_error) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < S; ++j)
          if (tab.stage(j)) acc = acc + k[j][n] * (tab.b2(j) - tab.b(j));
        e_curr[n] = acc * t_step;
      }
    }
  }

  // State::commit_step (state.rs:122-139)
  __device__ void commit_step() {
    hist.push(t_curr, p_curr, k, t_curr - t_prev);
    const double t_tail = t_prev - P.max_delay;
    hist.prune(t_tail);
    if (P.uses_disco) disco.prune(t_tail);
  }
  __device__ __forceinline__ void make_zero_step() {  // state.rs:141-144
    t_prev = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) p_prev[n] = p_curr[n];
  }
  __device__ __forceinline__ void undo_step() {  // state.rs:146-150
    t_curr = t_prev;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      p_curr[n] = p_prev[n];
      d_curr[n] = d_prev[n];
    }
  }

  // ---- stepsize (stepsize.rs) ----------------------------------------------
  __device__ __forceinline__ double stepsize_get() const {
    return P.stepsize_kind == DFB_STEP_AUTOMATIC ? h_auto : P.h;
  }
  // AutomaticStepsize::update (stepsize.rs:64-85); true = Rejected
  __device__ bool stepsize_update_rejects(double* h_before) {
    const dfb_auto_stepsize& A = P.automatic;
    *h_before = h_auto;
    double err = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const double ae = fabs(e_curr[n]);
      const double e = ae / (A.atol[n] + ae * A.rtol[n]);
      err = n == 0 ? e : fmax(err, e);
    }
#if DFB_ARITH == 0
    double factor = A.fac * dev_portable_root(1.0 / err, A.order + 1u);
#else
    double factor = A.fac * pow(1.0 / err, 1.0 / double(A.order + 1u));
#endif
    factor = dev_clamp(factor, A.fac_min, A.fac_max);
    h_auto *= factor;
    h_auto = dev_clamp(h_auto, A.stepsize_min, A.stepsize_max);
    return err >= 1.0;
  }

  // ---- EvalState over a locator's scalar function ---------------------------
  // which: 0 = eval_curr, 1 = eval_prev, 2 = eval_at(t)  (state_fn.rs:170-201)
  __device__ double loc_fn(int l, int which, double t) {
    const dfb_locator& L = P.loc[l];
    double y[N], dy[N];
    Ref v = which == 1 ? view_prev() : view_curr();
    if (which == 2) {
      state_eval<0>(t, y);
      state_eval<1>(t, dy);
      v = Ref{t, y, dy, t, y, &hist};
    }
    if (L.kind == DFB_LOC_PROPAGATION) {
      // Propagator as EvalState (loc.rs:424-432): delayed argument - tracked time
      const double arg = L.detector_id == 0 ? v.t - L.delay : Sys::detector(L.detector_id, v, prm);
      return arg - trk_t[l];
    }
    if (L.detection == DFB_DET_STEP) return 1.0;
    return Sys::detector(L.detector_id, v, prm);
  }

  // detection_method (loc.rs:127-143), Periodic::detect (:318-322), Propagator (:378-410)
  __device__ bool detect_base(int l) {
    const dfb_locator& L = P.loc[l];
    if (L.kind == DFB_LOC_PERIODIC) {
      const double prev = floor((t_prev - L.offset) / L.period);
      const double curr = floor((t_curr - L.offset) / L.period);
      return prev < curr;
    }
    if (L.kind == DFB_LOC_PROPAGATION) {
      const double keep = trk_t[l];
      trk_t[l] = 0.0;  // loc_fn subtracts the tracked time; we want the raw delayed argument
      const double prev = loc_fn(l, 1, 0.0);
      const double curr = loc_fn(l, 0, 0.0);
      trk_t[l] = keep;
      const int pp = disco.partition_point_linear(trk_index[l], prev);
      const int pc = disco.partition_point_linear(pp, curr);
      trk_index[l] = pp < pc ? pp : pc;
      if (pp != pc) {
        trk_t[l] = disco.t(trk_index[l]);
        trk_order[l] = (long long)disco.order(trk_index[l]) + L.smoothing_order;
        return true;
      }
      return false;
    }
    const int det = L.detection;
    if (det == DFB_DET_STEP) return true;
    const bool needs_prev = det == DFB_DET_ZERO || det == DFB_DET_ABOVE_ZERO ||
                            det == DFB_DET_BELOW_ZERO || det == DFB_DET_SWITCH ||
                            det == DFB_DET_SWITCH_TRUE || det == DFB_DET_SWITCH_FALSE;
    const double c = loc_fn(l, 0, 0.0);
    const double p = needs_prev ? loc_fn(l, 1, 0.0) : 0.0;
    const bool cb = c != 0.0, pb = p != 0.0;
    switch (det) {
      case DFB_DET_ZERO: return (c > 0.0 && p <= 0.0) || (c <= 0.0 && p > 0.0);
      case DFB_DET_ABOVE_ZERO: return c > 0.0 && p <= 0.0;
      case DFB_DET_BELOW_ZERO: return c < 0.0 && p >= 0.0;
      case DFB_DET_POSITIVE: return c >= 0.0;
      case DFB_DET_NEGATIVE: return c <= 0.0;
      case DFB_DET_SWITCH: return cb != pb;
      case DFB_DET_SWITCH_TRUE: return cb && !pb;
      case DFB_DET_SWITCH_FALSE: return !cb && pb;
      case DFB_DET_IS_TRUE: return cb;
      case DFB_DET_IS_FALSE: return !cb;
    }
    return false;
  }

  // location_method (loc.rs:200-285), Periodic::locate (:332-334)
  __device__ double locate_base(int l) {
    const dfb_locator& L = P.loc[l];
    if (L.kind == DFB_LOC_PERIODIC)
      return floor((t_curr - L.offset) / L.period) * L.period + L.offset;
    const int location = L.kind == DFB_LOC_PROPAGATION ? int(DFB_LOCATE_BISECTION) : L.location;
    if (location == DFB_LOCATE_STEP_BEGIN) return t_prev;
    if (location == DFB_LOCATE_STEP_END) return t_curr;
    if (location == DFB_LOCATE_STEP_MIDDLE) return 0.5 * (t_curr - t_prev);  // sic (loc.rs:202-204)
    const bool is_bool = location == DFB_LOCATE_BISECTION_BOOL;
    const double fc = is_bool ? 0.0 : loc_fn(l, 0, 0.0);
    const double fp = (is_bool || location == DFB_LOCATE_LERP) ? loc_fn(l, 1, 0.0) : 0.0;
    if (location == DFB_LOCATE_LERP) return (fc * t_prev - fp * t_curr) / (fc - fp);
    double lft = t_prev, rgt = t_curr;
    if (is_bool ? (fp != 0.0) : (fc < 0.0)) {
      const double tmp = lft;
      lft = rgt;
      rgt = tmp;
    }
    double w = fabs(rgt - lft);
    double w_prev = 2.0 * w;
    if (location == DFB_LOCATE_REGULA_FALSI) {
      while (w < w_prev) {
        w_prev = w;
        double fv[3];
        double m = 0.0;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
          if (q == 2) m = (fv[1] * lft - fv[0] * rgt) / (fv[1] - fv[0]);
          fv[q] = loc_fn(l, 2, q == 0 ? lft : (q == 1 ? rgt : m));
        }
        if (fv[2] < 0.0) rgt = m;
        else lft = m;
        w = fabs(rgt - lft);
      }
      return fmax(lft, rgt);
    }
    // Bisection (loc.rs:235-259) / BisectionBool (:210-234)
    double m = 0.5 * (lft + rgt);
    while (w < w_prev) {
      w_prev = w;
      const double f = loc_fn(l, 2, m);
      const bool go_left = is_bool ? !(f != 0.0) : (f < 0.0);
      if (go_left) lft = m;
      else rgt = m;
      m = 0.5 * (lft + rgt);
      w = fabs(rgt - lft);
    }
    return fmax(lft, rgt);
  }

  // Filter state machines (loc.rs:613-688)
  __device__ bool filter_pass(int l, double t) {
    const dfb_locator& L = P.loc[l];
    switch (L.filter) {
      case DFB_FILTER_EVERY:
        every_i[l] += 1;
        if (every_i[l] >= L.filter_n) {
          every_i[l] = 0;
          return true;
        }
        return false;
      case DFB_FILTER_SEPARATED_BY:
        if (t >= sep_prev[l] + L.filter_a) {
          sep_prev[l] = t;
          return true;
        }
        return false;
      case DFB_FILTER_IN_INTERVAL: return t >= L.filter_a && t < L.filter_b;
      case DFB_FILTER_ON_TIMES: {  // loc.rs:665-688: every_i counts detections, sep_prev walks the list
        const int i = every_i[l];
        const int pos = int(sep_prev[l]);
        every_i[l] = i + 1;
        if (pos < L.filter_n && i == L.filter_times[pos]) {
          sep_prev[l] = double(pos + 1);
          return true;
        }
        return false;
      }
    }
    return true;
  }

  // Detect phase of one locator, with DedupLocF (loc.rs:1002-1011) and the
  // filter wrappers (FilterAfterDetection :484-486, FilterLocated :568-576).
  __device__ bool detect_phase(int l) {
    const dfb_locator& L = P.loc[l];
    if (L.dedup && ((has_last_call >> l) & 1u) && !(t_prev > last_call[l])) return false;
    if (!detect_base(l)) return false;
    if (L.filter == DFB_FILTER_SEPARATED_BY || L.filter == DFB_FILTER_IN_INTERVAL) {
      // FilterLocated::detect = inner detect_and_locate + filter.eval_at(t);
      // LocCallback then calls locate() AGAIN (loc.rs:31-33, 733-746).
      return filter_pass(l, locate_base(l));
    }
    if (L.filter == DFB_FILTER_EVERY || L.filter == DFB_FILTER_ON_TIMES) return filter_pass(l, t_curr);
    return true;
  }

  // callback of locator l (LocCallback / DedupLocF::eval_mut loc.rs:1023-1026,
  // Propagation callback loc.rs:446-453)
  __device__ void fire_callback(int l) {
    const dfb_locator& L = P.loc[l];
    if (L.kind == DFB_LOC_PROPAGATION) {
      if (trk_order[l] < (long long)P.rk.order) {
        if (!disco.push(t_curr, int(trk_order[l]))) hist.status |= DFB_TRAJ_DISCO_OVERFLOW;
      }
      return;
    }
    if (L.dedup) {
      has_last_call |= 1u << l;
      last_call[l] = t_curr;
    }
    if (L.callback_id != DFB_CB_NONE) {
      RefMut m{&t_curr, p_curr, d_curr, t_prev, p_prev, &hist};
      Sys::callback(L.callback_id, m, prm, aux);
    }
  }

  // output policy standing in for on_step closures (solver.rs:290, 309)
  __device__ void on_step() {
    if (P.trace_every > 0) {
      if (on_step_calls % (unsigned long long)P.trace_every == 0 &&
          n_samples < uint32_t(P.trace_capacity)) {
        const size_t n_traj = P.n_traj;
        P.trace_t[size_t(n_samples) * n_traj + gid] = t_curr;
#pragma unroll
        for (int n = 0; n < N; ++n)
          P.trace_y[(size_t(n_samples) * N + n) * n_traj + gid] = p_curr[n];
        ++n_samples;
      }
      ++on_step_calls;
    }
  }

  __device__ void init() {
    const size_t n_traj = P.n_traj;
#pragma unroll
    for (int n = 0; n < Sys::NP; ++n) prm[n] = P.params[size_t(n) * n_traj + gid];
#pragma unroll
    for (int n = 0; n < NA; ++n) aux[n] = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) hist.y0[n] = P.y0[size_t(n) * n_traj + gid];
    hist.hist_k = Sys::NP >= 4 ? prm[Sys::NP >= 4 ? 3 : 0] : 0.0;
    // State::new (state.rs:52-81)
    t_curr = t_prev = P.t0;
    double p0[N];
    hist.template init_eval<0>(P.t0, p0);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      p_curr[n] = p_prev[n] = p0[n];
      d_curr[n] = d_prev[n] = e_curr[n] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < S; ++i)
#pragma unroll
      for (int n = 0; n < N; ++n) k[i][n] = 0.0;
    hist.push(P.t0, p0, nullptr, 0.0);
    if (P.uses_disco)
      for (int i = 0; i < P.n_disco; ++i) disco.push(P.disco_t[i], P.disco_order[i]);
    h_auto = P.automatic.stepsize;
    has_last_call = 0u;
    for (int l = 0; l < MAXL; ++l) {
      last_call[l] = 0.0;
      const bool on_times = l < P.n_loc && P.loc[l].filter == DFB_FILTER_ON_TIMES;
      every_i[l] = on_times ? 0 : (l < P.n_loc ? P.loc[l].filter_n - 1 : 0);
      sep_prev[l] = on_times ? 0.0 : -1.7976931348623157e308;  // T::min_value() / list position
      trk_t[l] = -1.7976931348623157e308;
      trk_order[l] = 0x7fffffffffffffffLL;    // usize::MAX
      trk_index[l] = 0;
    }
    n_steps = n_rejected = n_events = n_rhs = 0ull;
    n_samples = n_evlog = 0u;
    on_step_calls = 0ull;
    if (P.on_start_cb != DFB_CB_NONE) {  // events_on_start.eval_mut (solver.rs:289), before the first on_step
      RefMut m{&t_curr, p_curr, d_curr, t_prev, p_prev, &hist};
      Sys::callback(P.on_start_cb, m, prm, aux);
    }
  }

  __device__ void finish() {
    const size_t n_traj = P.n_traj;
    uint32_t status = hist.status;
    if (t_curr == dev_inf()) status |= DFB_TRAJ_STOPPED;
    bool finite = true;
#pragma unroll
    for (int n = 0; n < N; ++n) finite = finite && (fabs(p_curr[n]) < dev_inf());
    if (!finite) status |= DFB_TRAJ_NONFINITE;
    P.t_final[gid] = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) P.y_final[size_t(n) * n_traj + gid] = p_curr[n];
    if (P.aux_final) {
#pragma unroll
      for (int n = 0; n < Sys::NA; ++n) P.aux_final[size_t(n) * n_traj + gid] = aux[n];
    }
    P.steps[gid] = n_steps;
    P.rejected[gid] = n_rejected;
    P.events[gid] = n_events;
    P.rhs_evals[gid] = n_rhs;
    P.status[gid] = status;
    if (P.n_samples) P.n_samples[gid] = n_samples;
    if (P.n_events) P.n_events[gid] = n_evlog;
  }

  __device__ __forceinline__ int next_mode() {
    const bool more = t_curr < P.t1;  // branch-free (see ode_event.cuh)
    const bool limit = more && P.max_steps > 0 && n_steps >= (unsigned long long)P.max_steps;
    hist.status |= limit ? uint32_t(DFB_TRAJ_STEP_LIMIT) : 0u;
    return (more && !limit) ? MODE_STEP : MODE_DONE;
  }

  // ---- Solver::run (solver.rs:275-313) as a warp-friendly state machine -----
  // `valid` = this lane owns a trajectory.  All 32 lanes stay in the loop so
  // the warp votes are full-mask.
  __device__ void run(bool valid, int locate_batch_threshold, int locate_max_wait) {
    int mode = MODE_DONE;
    if (valid) {
      init();
      on_step();  // events_on_step at t_init (solver.rs:290)
      mode = next_mode();
    }
    double ev_time = 0.0;
    int ev_index = 0;
    uint32_t fired_mask = 0u;  // locators that detected on the parked step
    const bool adaptive = P.stepsize_kind == DFB_STEP_AUTOMATIC;
    int parked_iters = 0;  // warp-uniform: step rounds since the first lane parked

    while (true) {
      const unsigned stepping =
          __ballot_sync(0xffffffffu, mode == MODE_STEP || mode == MODE_EVENT_RESTEP);
      const unsigned parked = HAS_LOC ? __ballot_sync(0xffffffffu, mode == MODE_LOCATE) : 0u;
      if (stepping == 0u && parked == 0u) break;

      // (1) batched event location for the parked lanes: run the bisections once enough
      // lanes are parked, or the first parked lane has waited `locate_max_wait` step
      // rounds (bounds the idle time when events are rare), or nobody else can advance.
      if (HAS_LOC && parked != 0u &&
          (stepping == 0u || __popc(parked) >= locate_batch_threshold ||
           parked_iters >= locate_max_wait)) {
        parked_iters = 0;
        if (mode == MODE_LOCATE) {
          // HListLocateEarliest (loc.rs:821-835): strictly earliest, first index wins ties
          bool found = false;
          double best_t = 0.0;
          int best_i = 0;
#pragma unroll 1
          for (int l = 0; l < P.n_loc; ++l) {
            if ((fired_mask >> l) & 1u) {
              const double tl = locate_base(l);
              if (!found || tl < best_t) {
                found = true;
                best_t = tl;
                best_i = l;
              }
            }
          }
          if (hist.status & kPanicBits) {
            mode = MODE_DONE;  // a lookup inside a locator panicked upstream: the step is never committed
          } else if (found && best_t >= t_prev) {  // solver.rs:299-300
            undo_step();
            ev_time = best_t;
            ev_index = best_i;
            mode = MODE_EVENT_RESTEP;
          } else {
            commit_step();
            ++n_steps;
            on_step();
            mode = next_mode();
          }
        }
        continue;
      }

      if (HAS_LOC && parked != 0u) ++parked_iters;

      // (2) one make_step for every lane that can advance
      if (mode == MODE_STEP || mode == MODE_EVENT_RESTEP) {
        const bool restep = mode == MODE_EVENT_RESTEP;
        double h_before = 0.0;
        const double h_req = restep ? (ev_time - t_curr) : fmin(stepsize_get(), P.t1 - t_curr);
        make_step(h_req, adaptive);
        if (hist.status & kPanicBits) {
          mode = MODE_DONE;  // the reference panics here
        } else if (HAS_LOC && restep) {
          // solver.rs:304-309
          commit_step();
          make_zero_step();
          fire_callback(ev_index);
          ++n_events;
          if (P.event_capacity > 0) {
            if (n_evlog < uint32_t(P.event_capacity)) {
              P.ev_t[size_t(n_evlog) * P.n_traj + gid] = ev_time;
              P.ev_idx[size_t(n_evlog) * P.n_traj + gid] = ev_index;
            }
            ++n_evlog;
          }
          commit_step();
          ++n_steps;
          on_step();
          mode = next_mode();
        } else if (adaptive && stepsize_update_rejects(&h_before)) {
          ++n_rejected;
          undo_step();  // solver.rs:294-297
          // a rejection that leaves the controller's stepsize unchanged (clamped at
          // stepsize_range.start) repeats the same step from the same state: the reference would
          // retry forever; stop the trajectory instead (DFB_TRAJ_STEPSIZE_FLOOR)
          if (h_auto == h_before) {
            hist.status |= DFB_TRAJ_STEPSIZE_FLOOR;
            t_curr = dev_nan();
            mode = MODE_DONE;
          }
        } else {
          fired_mask = 0u;
          if (HAS_LOC) {
#pragma unroll 1
            for (int l = 0; l < P.n_loc; ++l)
              if (detect_phase(l)) fired_mask |= 1u << l;
          }
          if (hist.status & kPanicBits) {
            mode = MODE_DONE;  // a detector (or a FilterLocated locate) looked up a time the reference panics on
          } else if (fired_mask != 0u) {
            mode = MODE_LOCATE;
          } else {
            commit_step();
            ++n_steps;
            on_step();
            mode = next_mode();
          }
        }
      }
    }
    if (valid) finish();
  }
};

template <class Sys, int S, int I, bool HAS_LOC, bool RT = false, bool ODEHIST = false>
__global__ void __launch_bounds__(128)
    integrate_general_kernel(const __grid_constant__ DevProblem P, int locate_batch_threshold,
                             int locate_max_wait) {
  const size_t gid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = gid < P.n_traj;
  Integrator<Sys, S, I, HAS_LOC, RT, ODEHIST> it(P, valid ? gid : 0);
  it.run(valid, locate_batch_threshold, locate_max_wait);
}

}  // namespace DFB_NS
