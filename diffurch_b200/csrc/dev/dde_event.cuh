// Hot path #4: fixed-step explicit RK over a DDE/NDDE ensemble WITH located events and
// discontinuity propagation, FAST arithmetic (reference: Solver::run solver.rs:263-314;
// StateHistory state.rs:11-23, 122-139, 161-187; Detect/Locate, Propagator/Propagation,
// DedupLocF, filters loc.rs:127-143, 200-335, 344-454, 465-700, 795-883, 970-1052;
// examples/dde-relay-clamp.rs:24-35, tests/delay_propagation.rs:6-94).
//
// Same semantics as the general integrator (integrator.cuh) for this subset; built like the
// other hot-path kernels:
//
//  * State, every locator's bookkeeping and the propagation trackers live in registers
//    (locator loops unrolled over a compile-time locator count); the per-member discontinuity
//    list (state.rs:20) lives in shared memory;
//  * history = per-lane ring of COEFFICIENT records laid out [slot][warp][field][lane]: record r
//    holds (t_r, y_r, d_1..d_{I-1}) of the step that STARTS at t_r, y(t) = y_r + sum_j d_j (t-t_r)^j,
//    so a lookup touches one record and is an (I-1)-term Horner chain; event-truncated and
//    zero-length steps (solver.rs:304-308) make the time grid per lane, so the start time is stored;
//  * delayed-argument lookup: every lane caches the record its last lookup landed in
//    (index, start and end time, column pointer).  Delayed arguments move forward through the
//    history, so most lookups hit the cache without touching the time column; a miss walks from
//    the cached index (one or two probes) with a bisection fallback — the same index the
//    reference's partition_point (state.rs:165) returns.  A lane's ring is its own column of the
//    warp's block, so the 32 lanes of a warp read one coalesced row per field whenever they sit on
//    the same slot (they drift apart only by the records their events added);
//  * one make_step call site, warp-batched event location and dynamic member assignment, as in
//    ode_event.cuh; the step's interpolant is pre-combined once per step — the same
//    coefficients serve the bisection (State::eval of the step in flight) and the ring record.
#pragma once
#include "common.cuh"
#include "history_fn.cuh"
#include "integrator.cuh"
#include "systems.cuh"

namespace DFB_NS {

struct DdeEventArgs {
  double a[DFB_MAX_STAGES * DFB_MAX_STAGES];
  double b[DFB_MAX_STAGES];
  double c[DFB_MAX_STAGES];
  double bi[DFB_MAX_STAGES * DFB_MAX_INTERP];
  double t0, t1, h, max_delay;
  long long max_steps;
  int32_t history_id;
  int32_t ring_capacity;  // power of two
  int32_t rk_order;       // ButcherTableu::order: propagated discontinuities of lower order are kept (loc.rs:446-453)
  int32_t n_disco;
  double disco_t[DFB_MAX_DISCO];
  int32_t disco_order[DFB_MAX_DISCO];
  size_t n_traj;
  const double* y0;
  const double* params;
  double* t_final;
  double* y_final;
  double* aux_final;
  unsigned long long* steps;
  unsigned long long* events;
  unsigned long long* rhs_evals;
  unsigned long long* locate_iters;  // location-loop iterations per member (may be null)
  uint32_t* status;
  int32_t n_loc, event_capacity;
  dfb_locator loc[DFB_MAX_LOCATORS];
  double* ev_t;
  int32_t* ev_idx;
  uint32_t* n_events;
  int32_t trace_every, trace_capacity;
  double* trace_t;
  double* trace_y;
  uint32_t* n_samples;
  unsigned long long* queue;
  int32_t locate_batch, locate_wait, inner_steps, on_start_cb;
  double* ring;            // [cap][ring_warps][F][32], F = 1 + N I
  unsigned ring_warps;     // warps the ring was allocated for (grid warps <= ring_warps)
  int32_t prefetch_mode;   // 0 none, 1 prefetch.global.L1, 2 prefetch.global.L2 (DFB_DDE_EV_PREFETCH_MODE, tuning)
};

// Partition point of one lane's time column (first absolute record index in [base, n_rec] with
// t_r > t), started from the guess g: forward while t_g <= t, backward while t_{g-1} > t, at most 8
// probes, then bisection of what is left.
static __device__ __noinline__ int ring_partition_point(const double* col, unsigned slot_stride, unsigned mask,
                                                        int base, int n_rec, int g, double t) {
  auto T = [&](int r) { return col[(unsigned(r) & mask) * slot_stride]; };
  g = g < base ? base : (g > n_rec ? n_rec : g);
  int probes = 0;
  bool settled = false;
  int lo = base, hi = n_rec;
  while (probes < 8) {
    if (g < n_rec && T(g) <= t) {
      lo = ++g;
      ++probes;
    } else {
      hi = g;
      break;
    }
  }
  if (probes < 8) {
    while (probes < 8) {
      if (g > lo && !(T(g - 1) <= t)) {
        hi = --g;
        ++probes;
      } else {
        settled = true;
        break;
      }
    }
  }
  if (!settled) {
    while (lo < hi) {
      const int mid = lo + (hi - lo) / 2;
      if (T(mid) <= t) lo = mid + 1;
      else hi = mid;
    }
    g = lo;
  }
  return g;
}

// Everything a lookup does off its fast path — the history function (t <= t_init,
// initial_condition.rs:15-50, which drags libm's sin/cos slow paths along), the search of the time
// column, the reference's two panics (state.rs:166-177) and the evaluation of a record outside the
// window — is ONE out-of-line copy per kernel that takes and returns plain values.  Inlined at each
// of the ~20 lookup sites it made the first build of this kernel 250 KB of code (instruction-fetch
// stalls on every taken branch) and forced register copies of the window state around every lookup.
template <int N>
struct LookupValue {
  double v[N];
  uint32_t status;
};
template <int D, int N, int I>
static __device__ __noinline__ LookupValue<N> ring_lookup_slow(const DdeEventArgs* A, const double* col,
                                                               unsigned slot_stride, unsigned mask, int base,
                                                               int n_rec, int guess, size_t gid, double t) {
  LookupValue<N> r;
  r.status = 0u;
  if (t <= A->t0) {
    double y0[N];
#pragma unroll
    for (int n = 0; n < N; ++n) y0[n] = A->y0[size_t(n) * A->n_traj + gid];
    const double hk = A->history_id == DFB_HIST_CONSTANT ? 0.0 : A->params[size_t(3) * A->n_traj + gid];
    init_condition_eval<D, N>(A->history_id, t, y0, hk, r.status, r.v);
    return r;
  }
  const int g = ring_partition_point(col, slot_stride, mask, base, n_rec, guess, t);
  if (g == base || g == n_rec) {
    r.status = (g == base) ? DFB_TRAJ_HISTORY_DELETED : DFB_TRAJ_HISTORY_NOT_READY;
#pragma unroll
    for (int n = 0; n < N; ++n) r.v[n] = dev_nan();
    return r;
  }
  const double* p = col + (unsigned(g - 1) & mask) * slot_stride;
  const double s = t - p[0];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    if (I == 1) {
      r.v[n] = D == 0 ? p[32 * (1 + n)] : 0.0;
    } else if (D == 0) {
      double acc = p[32 * (1 + N * (I - 1) + n)];
#pragma unroll
      for (int j = I - 2; j >= 1; --j) acc = fma(acc, s, p[32 * (1 + N * j + n)]);
      r.v[n] = fma(acc, s, p[32 * (1 + n)]);
    } else {
      double acc = double(I - 1) * p[32 * (1 + N * (I - 1) + n)];
#pragma unroll
      for (int j = I - 2; j >= 1; --j) acc = fma(acc, s, double(j) * p[32 * (1 + N * j + n)]);
      r.v[n] = acc;
    }
  }
  return r;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

#ifndef DFB_DDE_EV_PRE
#define DFB_DDE_EV_PRE 0  // delayed values of a step looked up ahead of its stage chain: measured slower (see PreHist)
#endif
#ifndef DFB_DDE_EV_PREFETCH
#define DFB_DDE_EV_PREFETCH 1
#endif

// History of one lane over its column of the warp's ring block.
//
// Window.  Two consecutive records A = [a0, a1) and B = [a1, a2) are "open": a lookup inside
// [a0, a2) is two compares, a select of the record pointer and a Horner chain, and it MODIFIES
// NOTHING — so the lookups of a step are independent instructions the scheduler can overlap with the
// stage arithmetic (a cache that moves inside the lookup puts a branch and a block of register copies
// at every site).  The window moves once per committed step (`slide`), to the record that holds the
// last lookup of the step just taken: delayed arguments move forward, so that is the earliest time the
// next step can ask for.  The end time of the record after B (a3) is loaded one slide ahead and B's
// successor is prefetched then, so no lookup waits on a first-touch HBM read.  Anything outside the
// window takes ring_lookup_slow — same result, found by search.
template <int N, int I>
struct RingHist {
  static constexpr int F = 1 + N * I;  // doubles per record: t, y[N], d_j[N] (j = 1 .. I-1)
  const DdeEventArgs* A;
  double* col;           // this lane's column of slot 0
  unsigned slot_stride;  // doubles between two slots
  unsigned mask;         // ring_capacity - 1
  size_t gid;
  int n_rec;             // records pushed; record n_rec-1 is the open end (last committed time and state)
  int base;              // absolute index of the oldest retained record (t_deque[0])
  double t_base1;        // T(base + 1), +inf while fewer than two records are retained
  double t_base2;        // T(base + 2), loaded one pop ahead so that the prune never waits on its own load
  int w;                 // window: records w (A) and w + 1 (B)
  double a0, a1, a2, a3; // their time spans and T(w + 3); -inf = that time is not committed yet
  const double *pA, *pB;
  double t_look;         // time of the latest lookup (dead-store eliminated down to the last one of a block)
  uint32_t status;

  __device__ __forceinline__ double* rec(int r) const { return col + (unsigned(r) & mask) * slot_stride; }
  __device__ __forceinline__ double T(int r) const { return *rec(r); }

  __device__ __forceinline__ void close_window() {
    a0 = dev_inf();
    a1 = a2 = a3 = -dev_inf();
  }
  __device__ __forceinline__ void reset(size_t gid_) {
    gid = gid_;
    n_rec = 0;
    base = 0;
    t_base1 = t_base2 = dev_inf();
    w = 0;
    close_window();
    pA = pB = col;
    t_look = -dev_inf();
    status = 0u;
  }

  __device__ __forceinline__ void prefetch_record(int r) {
#if DFB_DDE_EV_PREFETCH
    const double* q = rec(r);
    if (A->prefetch_mode == 1) {
#pragma unroll
      for (int f = 1; f < F; ++f) prefetch_l1(q + 32 * f);
    } else if (A->prefetch_mode == 2) {
#pragma unroll
      for (int f = 1; f < F; ++f) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 32 * f));
    } else if (A->prefetch_mode == 3) {
#pragma unroll
      for (int f = 1; f < F; ++f) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(q + 32 * f));
    }
#endif
  }

  // Called once per committed step with the time of the last lookup that step made.
  __device__ __forceinline__ void slide(double t_hint) {
    if (!(t_hint > A->t0)) return;  // still inside the history function
    if (!(a0 <= t_hint)) {          // closed (start-up, pruned under us) or behind the hint: open it by search
      const int g = ring_partition_point(col, slot_stride, mask, base, n_rec, w + 1, t_hint);
      if (g == base || g == n_rec) return;
      w = g - 1;
      pA = rec(w);
      pB = rec(w + 1);
      a0 = *pA;
      a1 = *pB;
      a2 = (w + 2 < n_rec) ? T(w + 2) : -dev_inf();
      a3 = (w + 3 < n_rec) ? T(w + 3) : -dev_inf();
      return;
    }
    // times that were not committed when the window was set may be by now
    if (a2 == -dev_inf() && w + 2 < n_rec) a2 = T(w + 2);
    if (a3 == -dev_inf() && w + 3 < n_rec) a3 = T(w + 3);
#pragma unroll 1
    while (t_hint >= a1 && a2 != -dev_inf()) {  // A is behind the hint and B is complete: move on by one
      ++w;
      a0 = a1;
      a1 = a2;
      a2 = a3;
      pA = pB;
      pB = rec(w + 1);
      a3 = (w + 3 < n_rec) ? T(w + 3) : -dev_inf();
      if (w + 2 < n_rec) prefetch_record(w + 2);  // B of the next slide
    }
  }

  template <int D>
  __device__ __forceinline__ void eval(double t, double* out) {
    t_look = t;
    const bool fast = (t >= a0) & (t < a2) & (t > A->t0);
    if (!fast) {
      const LookupValue<N> r = ring_lookup_slow<D, N, I>(A, col, slot_stride, mask, base, n_rec, w + 1, gid, t);
      status |= r.status;
#pragma unroll
      for (int n = 0; n < N; ++n) out[n] = r.v[n];
      return;
    }
    const bool second = t >= a1;
    const double* p = second ? pB : pA;
    const double s = t - (second ? a1 : a0);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      if (I == 1) {
        out[n] = D == 0 ? p[32 * (1 + n)] : 0.0;
      } else if (D == 0) {
        double acc = p[32 * (1 + N * (I - 1) + n)];
#pragma unroll
        for (int j = I - 2; j >= 1; --j) acc = fma(acc, s, p[32 * (1 + N * j + n)]);
        out[n] = fma(acc, s, p[32 * (1 + n)]);
      } else {
        double acc = double(I - 1) * p[32 * (1 + N * (I - 1) + n)];
#pragma unroll
        for (int j = I - 2; j >= 1; --j) acc = fma(acc, s, double(j) * p[32 * (1 + N * j + n)]);
        out[n] = acc;
      }
    }
  }

  // push of commit_step (state.rs:123-125).  `cf(j, n)` = coefficients of the step that ends at t; they go
  // into the record the step STARTED at, then (t, y) opens the next one.
  template <class Cf>
  __device__ __forceinline__ void push(double t, const double* y, bool have_cf, bool cf_valid, Cf cf) {
    if (have_cf && n_rec > 0) {
      double* p = rec(n_rec - 1);
#pragma unroll
      for (int j = 1; j < I; ++j)
#pragma unroll
        for (int n = 0; n < N; ++n) __stcs(p + 32 * (1 + N * j + n), cf_valid ? cf(j - 1, n) : dev_nan());
    }
    if (n_rec - base == int(mask) + 1) {  // library limit, not a reference behaviour
      status |= DFB_TRAJ_RING_OVERFLOW;
      pop_front();
      if (w < base) close_window();
    }
    double* p = rec(n_rec);
    __stcs(p, t);
#pragma unroll
    for (int n = 0; n < N; ++n) __stcs(p + 32 * (1 + n), y[n]);
    ++n_rec;
    if (n_rec - base == 2) t_base1 = t;
    if (n_rec - base == 3) t_base2 = t;
  }
  __device__ __forceinline__ void pop_front() {
    ++base;
    t_base1 = t_base2;
    t_base2 = (n_rec - base > 2) ? T(base + 2) : dev_inf();
  }

  // prune of commit_step (state.rs:126-133): pop while t_deque[1] < t_tail
  __device__ __forceinline__ void prune(double t_tail) {
    while (n_rec - base > 1 && t_base1 < t_tail) pop_front();
    if (w < base) close_window();  // a lookup of a pruned time must take the slow path (it panics upstream)
  }
};

// Per-member discontinuity list (state.rs:20, 134-138) in shared memory: [DFB_DISCO_CAP][BLOCK], followed
// by two rows per locator that cache the GAP of the list the locator's delayed argument sits in
// (lo <= argument < hi with no list entry in between): while both ends of a step stay in the gap nothing
// can be detected (loc.rs:378-410 finds equal partition points), and the per-step walk of the list shrinks
// to four compares.  Any change of the list closes every gap.
template <int NLOC>
struct SmemDisco {
  double* t_col;   // this lane's column
  int32_t* o_col;
  int stride;      // BLOCK
  int head, len;
  __device__ __forceinline__ int slot(int r) const { return ((head + r) & (DFB_DISCO_CAP - 1)) * stride; }
  __device__ __forceinline__ double t(int r) const { return t_col[slot(r)]; }
  __device__ __forceinline__ int order(int r) const { return o_col[slot(r)]; }
  __device__ __forceinline__ double& gap_lo(int l) { return t_col[(DFB_DISCO_CAP + 2 * l) * stride]; }
  __device__ __forceinline__ double& gap_hi(int l) { return t_col[(DFB_DISCO_CAP + 2 * l + 1) * stride]; }
  __device__ __forceinline__ void close_gaps() {
#pragma unroll
    for (int l = 0; l < NLOC; ++l) {
      gap_lo(l) = dev_inf();
      gap_hi(l) = -dev_inf();
    }
  }
  __device__ __forceinline__ bool push(double tt, int ord) {
    if (len == DFB_DISCO_CAP) return false;
    const int s = slot(len);
    t_col[s] = tt;
    o_col[s] = ord;
    ++len;
    close_gaps();
    return true;
  }
  __device__ __forceinline__ void prune(double t_tail) {
    if (len > 0 && t(0) < t_tail) {
      do {
        head = (head + 1) & (DFB_DISCO_CAP - 1);
        --len;
      } while (len > 0 && t(0) < t_tail);
      close_gaps();
    }
  }
  // util::partition_point_linear (src/util/partition_point.rs:26-52), pred = t_i <= x
  __device__ __forceinline__ int partition_point_linear(int start, double x) const {
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

// SIMPLE = every locator is either a propagated-discontinuity locator (with_const_delay /
// with_delayed_argument) or what Locator::{zero, above_zero, below_zero} builds (state function,
// sign-change detection, Bisection), all unfiltered: the shape of examples/dde-relay-clamp.rs and
// tests/delay_propagation.rs.  Known at compile time, the Periodic / filter / boolean-detection /
// other-location code disappears from the kernel (it is what the instruction cache has to hold).
// KINDS >= 0 (SIMPLE only): bit l set = locator l is a propagated-discontinuity locator, known at compile
// time, so the per-step detection code has no run-time dispatch on the locator kind; KINDS < 0 = read it
// from the descriptor.
template <class Sys, int S_, int I_, int NLOC, bool HAS_PROP, bool SIMPLE, int KINDS = -1>
struct DdeEventStepper {
  static constexpr int S = S_, I = I_;
  static constexpr int N = Sys::N;
  static constexpr int NP = Sys::NP > 0 ? Sys::NP : 1;
  static constexpr int NA = Sys::NA > 0 ? Sys::NA : 1;
  static constexpr int NCF = I > 1 ? I - 1 : 1;
  static_assert(!Sys::reads_d, "views of this kernel carry the fresh k[0] where the reference carries the stale d");
  using Hist = RingHist<N, I>;
  using Ref = StateRef<N, Hist>;
  using RefMut = StateRefMut<N, Hist>;

  const DdeEventArgs& A;
  Hist hist;
  SmemDisco<NLOC> disco;

  __device__ __forceinline__ static bool is_prop(const dfb_locator& L, int l) {
    if (!HAS_PROP) return false;
    if (SIMPLE && KINDS >= 0) return (KINDS >> l) & 1;
    return L.kind == DFB_LOC_PROPAGATION;
  }
  double t_curr, t_prev;
  double p_curr[N], p_prev[N], d_curr[N];
  double k[S][N];
  // Interpolant of the step in flight, y(t) = p_prev + sum_j cf(j-1) (t - t_prev)^j.  The stages are dead
  // once a step is taken (the next step recomputes them; only k[0] = d_prev is needed for the undo), so
  // the coefficients overwrite k[1 .. I-1] where the tableau has that many stages
  static constexpr bool CF_IN_K = I <= S;
  double cf_own[CF_IN_K ? 1 : NCF][N];
  __device__ __forceinline__ double& cfv(int j, int n) { return CF_IN_K ? k[CF_IN_K ? j + 1 : 0][n] : cf_own[CF_IN_K ? 0 : j][n]; }
  double t_slide;  // last lookup time of the latest make_step: where the history window moves at commit
  double prm[NP], aux[NA];
  double last_call[NLOC], sep_prev[NLOC], trk_t[NLOC];
  int trk_order[NLOC];  // usize in the reference; INT_MAX stands for usize::MAX (never incremented)
  int every_i[NLOC], trk_index[NLOC];
  // detector values of the step end, reused as the next step's eval_prev when the detectors are pure
  // functions of (t, p, d, history): the same inputs give the same bits, and the lookups keep moving
  // forward through the history (the eval_prev lookup would step back by h every step)
  double det_cache[NLOC];
  uint32_t det_valid, det_valid_prev;
  uint32_t has_last_call;
  unsigned n_steps, n_events, n_evlog, n_samples, n_iters, on_step_calls;
  unsigned n_make, n_k0;  // make_step calls and k[0] evaluations: rhs_evals = S n_make + n_k0
  unsigned step_limit;    // max_steps as a 32-bit bound (0xFFFFFFFF = none), set once per kernel

  __device__ explicit DdeEventStepper(const DdeEventArgs& A_) : A(A_) {}

  __device__ __forceinline__ Ref view_curr() { return Ref{t_curr, p_curr, d_curr, t_prev, p_prev, &hist}; }
  __device__ __forceinline__ Ref view_prev() { return Ref{t_prev, p_prev, k[0], t_prev, p_prev, &hist}; }

  // History of the views handed to the right-hand side when the system's one delay is constant
  // (Sys::has_const_delay): the delayed values of the whole step are looked up BEFORE its stage chain — their
  // times t_prev + c_i h - tau do not depend on the state — so the S ring reads (L1 / L2 / HBM latency each)
  // overlap instead of standing one after the other inside the stages.  The functor still calls
  // s.p_at(s.t - tau): it is handed the value computed for that very time and the ring for any other.
  // EXPERIMENT, off (DFB_DDE_EV_PRE=0): what pays for a lone warp of the constant-delay kernel (dde_fixed.cuh)
  // costs here — relay-clamp 2^16 / 2^18 members 9.47 -> 10.07 / 34.2 -> 38.4 ms: with 16 warps per SM the
  // ring reads are already overlapped by other warps, and the S values held across the stage chain push a
  // kernel that spills at 128 registers further (profiles/r2_dde_event_tuning.txt).
  struct PreHist {
    Hist* h;
    double t_key;
    const double* val;
    template <int D>
    __device__ __forceinline__ void eval(double t, double* out) {
      if (D == 0 && t == t_key) {
#pragma unroll
        for (int n = 0; n < N; ++n) out[n] = val[n];
      } else {
        h->template eval<D>(t, out);
      }
    }
  };
  static constexpr bool PRE = (DFB_DDE_EV_PRE != 0) && Sys::has_const_delay;

  template <bool P = PRE>
  __device__ __forceinline__ void rhs_stage(double t_key, const double* val, double* out) {
    if constexpr (P) {
      PreHist ph{&hist, t_key, val};
      Sys::rhs(StateRef<N, PreHist>{t_curr, p_curr, d_curr, t_prev, p_prev, &ph}, prm, out);
    } else {
      Sys::rhs(view_curr(), prm, out);
    }
  }

  // ---- State::make_step (state.rs:94-120); the caller guarantees d_curr is current --------
  __device__ __forceinline__ void make_step(double t_step) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      k[0][n] = d_curr[n];
      p_prev[n] = p_curr[n];
    }
    t_prev = t_curr;
    double pre_t[PRE ? S : 1], pre_p[PRE ? S : 1][N];  // slot i-1 = stage i, slot S-1 = the step end
    if constexpr (PRE) {
      const double tau = Sys::const_delay(prm);
#pragma unroll
      for (int i = 1; i <= S; ++i) {
        const double tt = (i < S ? t_prev + A.c[i] * t_step : t_prev + t_step) - tau;
        pre_t[PRE ? i - 1 : 0] = tt;
        hist.template eval<0>(tt, pre_p[PRE ? i - 1 : 0]);
      }
    }
#pragma unroll
    for (int i = 1; i < S; ++i) {
      t_curr = t_prev + A.c[i] * t_step;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < i; ++j) acc = acc + k[j][n] * A.a[i * DFB_MAX_STAGES + j];
        p_curr[n] = p_prev[n] + acc * t_step;
      }
      rhs_stage(pre_t[PRE ? i - 1 : 0], pre_p[PRE ? i - 1 : 0], k[i]);
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < S; ++j) acc = acc + k[j][n] * A.b[j];
      p_curr[n] = p_prev[n] + acc * t_step;
    }
    t_curr = t_prev + t_step;
    rhs_stage(pre_t[PRE ? S - 1 : 0], pre_p[PRE ? S - 1 : 0], d_curr);
    ++n_make;
    // pre-combined interpolant: cf_j = sum_i k_i bi[i][j] / t_step^(j-1)  (needs b_i(0) = 0: host-checked)
    t_slide = hist.t_look;
    const double inv = 1.0 / t_step;
    double scale = 1.0;
    double c[NCF][N];
#pragma unroll
    for (int j = 1; j < I; ++j) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < S; ++i) acc = fma(k[i][n], A.bi[i * DFB_MAX_INTERP + j], acc);
        c[j - 1][n] = acc * scale;
      }
      scale *= inv;
    }
#pragma unroll
    for (int j = 1; j < I; ++j)
#pragma unroll
      for (int n = 0; n < N; ++n) cfv(j - 1, n) = c[j - 1][n];
  }

  // State::eval::<D> (state.rs:83-92): the step in flight, else the history
  template <int D>
  __device__ __forceinline__ void state_eval(double t, double* out) {
    if (t >= t_prev && t <= t_curr) {
      const double s = t - t_prev;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        if (I == 1) {
          out[n] = D == 0 ? p_prev[n] : 0.0;
        } else if (D == 0) {
          double acc = cfv(I - 2, n);
#pragma unroll
          for (int j = I - 3; j >= 0; --j) acc = fma(acc, s, cfv(j, n));
          out[n] = fma(acc, s, p_prev[n]);
        } else {
          double acc = double(I - 1) * cfv(I - 2, n);
#pragma unroll
          for (int j = I - 3; j >= 0; --j) acc = fma(acc, s, double(j + 1) * cfv(j, n));
          out[n] = acc;
        }
      }
    } else {
      hist.template eval<D>(t, out);
    }
  }

  // State::commit_step (state.rs:122-139)
  __device__ __forceinline__ void commit_step(bool cf_valid) {
    hist.push(t_curr, p_curr, true, cf_valid, [&](int j, int n) { return cfv(j, n); });
    const double t_tail = t_prev - A.max_delay;
    hist.prune(t_tail);
    if (HAS_PROP) disco.prune(t_tail);
    hist.slide(t_slide);
  }

  // ---- EvalState over a locator's scalar function; which: 0 = curr, 1 = prev, 2 = at(t) -----
  // RAW: Propagator wants the delayed argument itself, not minus the tracked time (loc.rs:378-432)
  template <bool RAW = false>
  __device__ __forceinline__ double loc_fn(const dfb_locator& L, int l, int which, double t) {
    if (!SIMPLE && L.kind == DFB_LOC_STATEFN && L.detection == DFB_DET_STEP) return 1.0;
    double y[N], dy[N];
    Ref v = which == 1 ? view_prev() : view_curr();
    if (which == 2) {
      state_eval<0>(t, y);
      if (Sys::reads_d) state_eval<1>(t, dy);
      v = Ref{t, y, Sys::reads_d ? dy : y, t, y, &hist};
    }
    if (is_prop(L, l)) {
      const double arg = L.detector_id == 0 ? v.t - L.delay : Sys::detector(L.detector_id, v, prm);
      return RAW ? arg : arg - trk_t[l];
    }
    return Sys::detector(L.detector_id, v, prm);
  }

  // eval_prev of a state-function locator (state_fn.rs:181-188)
  __device__ __forceinline__ double loc_fn_prev(const dfb_locator& L, int l) {
    if (Sys::pure_detectors && ((det_valid_prev >> l) & 1u)) return det_cache[l];
    return loc_fn(L, l, 1, 0.0);
  }
  __device__ __forceinline__ double loc_fn_curr(const dfb_locator& L, int l) {
    const double c = loc_fn(L, l, 0, 0.0);
    if (Sys::pure_detectors) {
      det_cache[l] = c;
      det_valid |= 1u << l;
    }
    return c;
  }

  // detection_method (loc.rs:127-143), Periodic::detect (:318-322), Propagator (:378-410)
  __device__ __forceinline__ bool detect_base(const dfb_locator& L, int l) {
    if (!SIMPLE && L.kind == DFB_LOC_PERIODIC) {
      const double prev = floor((t_prev - L.offset) / L.period);
      const double curr = floor((t_curr - L.offset) / L.period);
      return prev < curr;
    }
    if (is_prop(L, l)) {
      const double prev = loc_fn<true>(L, l, 1, 0.0);
      const double curr = loc_fn<true>(L, l, 0, 0.0);
      const double lo = disco.gap_lo(l), hi = disco.gap_hi(l);
      if ((prev >= lo) & (prev < hi) & (curr >= lo) & (curr < hi)) return false;  // both in one gap of the list
      const int pp = disco.partition_point_linear(trk_index[l], prev);
      const int pc = disco.partition_point_linear(pp, curr);
      trk_index[l] = pp < pc ? pp : pc;
      if (pp != pc) {
        trk_t[l] = disco.t(trk_index[l]);
        trk_order[l] = disco.order(trk_index[l]) + L.smoothing_order;
        return true;
      }
      disco.gap_lo(l) = pp > 0 ? disco.t(pp - 1) : -dev_inf();
      disco.gap_hi(l) = pp < disco.len ? disco.t(pp) : dev_inf();
      return false;
    }
    const int det = L.detection;
    if (SIMPLE) {
      const double p = loc_fn_prev(L, l);  // before loc_fn_curr overwrites the cached value
      const double c = loc_fn_curr(L, l);
      const bool up = c > 0.0 && p <= 0.0;      // AboveZero (loc.rs:130-132)
      const bool down_z = c <= 0.0 && p > 0.0;  // the falling half of Zero (loc.rs:127-129)
      const bool down = c < 0.0 && p >= 0.0;    // BelowZero (loc.rs:133-135)
      return det == DFB_DET_ZERO ? (up || down_z) : (det == DFB_DET_ABOVE_ZERO ? up : down);
    }
    if (det == DFB_DET_STEP) return true;
    const bool needs_prev = det == DFB_DET_ZERO || det == DFB_DET_ABOVE_ZERO || det == DFB_DET_BELOW_ZERO ||
                            det == DFB_DET_SWITCH || det == DFB_DET_SWITCH_TRUE || det == DFB_DET_SWITCH_FALSE;
    const double p = needs_prev ? loc_fn_prev(L, l) : 0.0;  // before loc_fn_curr overwrites the cached value
    const double c = loc_fn_curr(L, l);
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

  // location_method (loc.rs:200-285), Periodic::locate (:332-334); Propagator locates by Bisection (:412-422)
  __device__ __forceinline__ double locate_base(const dfb_locator& L, int l) {
    if (SIMPLE) {  // Bisection (loc.rs:235-259)
      double lft = t_prev, rgt = t_curr;
      if (loc_fn(L, l, 0, 0.0) < 0.0) {
        lft = t_curr;
        rgt = t_prev;
      }
      double w = fabs(rgt - lft);
      double w_prev = 2.0 * w;
      double m = 0.5 * (lft + rgt);
#pragma unroll 1
      while (w < w_prev) {
        w_prev = w;
        if (loc_fn(L, l, 2, m) < 0.0) lft = m;
        else rgt = m;
        m = 0.5 * (lft + rgt);
        w = fabs(rgt - lft);
        ++n_iters;
      }
      return fmax(lft, rgt);
    }
    if (L.kind == DFB_LOC_PERIODIC) return floor((t_curr - L.offset) / L.period) * L.period + L.offset;
    const int location = is_prop(L, l) ? int(DFB_LOCATE_BISECTION) : L.location;
    if (location == DFB_LOCATE_STEP_BEGIN) return t_prev;
    if (location == DFB_LOCATE_STEP_END) return t_curr;
    if (location == DFB_LOCATE_STEP_MIDDLE) return 0.5 * (t_curr - t_prev);  // sic (loc.rs:202-204)
    const bool is_bool = location == DFB_LOCATE_BISECTION_BOOL;
    const double fc = is_bool ? 0.0 : loc_fn(L, l, 0, 0.0);
    const double fp = (is_bool || location == DFB_LOCATE_LERP) ? loc_fn(L, l, 1, 0.0) : 0.0;
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
          fv[q] = loc_fn(L, l, 2, q == 0 ? lft : (q == 1 ? rgt : m));
        }
        if (fv[2] < 0.0) rgt = m;
        else lft = m;
        w = fabs(rgt - lft);
        ++n_iters;
      }
      return fmax(lft, rgt);
    }
    // Bisection (loc.rs:235-259) / BisectionBool (:210-234)
    double m = 0.5 * (lft + rgt);
#pragma unroll 1
    while (w < w_prev) {
      w_prev = w;
      const double f = loc_fn(L, l, 2, m);
      const bool go_left = is_bool ? !(f != 0.0) : (f < 0.0);
      if (go_left) lft = m;
      else rgt = m;
      m = 0.5 * (lft + rgt);
      w = fabs(rgt - lft);
      ++n_iters;
    }
    return fmax(lft, rgt);
  }

  // Filter state machines (loc.rs:613-688); l is a compile-time index after unrolling
  __device__ __forceinline__ bool filter_pass(const dfb_locator& L, int l, double t) {
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
      case DFB_FILTER_ON_TIMES: {
        const int i = every_i[l];
        const int pos = int(sep_prev[l]);
        every_i[l] = i + 1;
        bool hit = false;
#pragma unroll
        for (int q = 0; q < DFB_MAX_FILTER_TIMES; ++q)
          if (q == pos && q < L.filter_n && i == L.filter_times[q]) hit = true;
        if (hit) sep_prev[l] = double(pos + 1);
        return hit;
      }
    }
    return true;
  }

  // DedupLocF (loc.rs:1002-1011) + filter wrappers (FilterAfterDetection :484-486, FilterLocated :568-576)
  __device__ __forceinline__ bool detect_phase(const dfb_locator& L, int l) {
    if (L.dedup && ((has_last_call >> l) & 1u) && !(t_prev > last_call[l])) return false;
    if (!detect_base(L, l)) return false;
    if (SIMPLE) return true;
    if (L.filter == DFB_FILTER_SEPARATED_BY || L.filter == DFB_FILTER_IN_INTERVAL)
      return filter_pass(L, l, locate_base(L, l));
    if (L.filter == DFB_FILTER_EVERY || L.filter == DFB_FILTER_ON_TIMES) return filter_pass(L, l, t_curr);
    return true;
  }

  // LocCallback / DedupLocF::eval_mut (loc.rs:1023-1026), Propagation callback (loc.rs:446-453)
  __device__ __forceinline__ void fire_callback(const dfb_locator& L, int l) {
    if (is_prop(L, l)) {
      if (trk_order[l] < A.rk_order) {
        if (!disco.push(t_curr, trk_order[l])) hist.status |= DFB_TRAJ_DISCO_OVERFLOW;
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

  __device__ __forceinline__ void on_step() {
    if (A.trace_every > 0) {
      if (on_step_calls % unsigned(A.trace_every) == 0u && n_samples < uint32_t(A.trace_capacity)) {
        const size_t n_traj = A.n_traj;
        A.trace_t[size_t(n_samples) * n_traj + hist.gid] = t_curr;
#pragma unroll
        for (int n = 0; n < N; ++n) A.trace_y[(size_t(n_samples) * N + n) * n_traj + hist.gid] = p_curr[n];
        ++n_samples;
      }
      ++on_step_calls;
    }
  }

  __device__ __forceinline__ int next_mode() {
    const bool more = t_curr < A.t1;
    const bool limit = more && n_steps >= step_limit;
    hist.status |= limit ? uint32_t(DFB_TRAJ_STEP_LIMIT) : 0u;
    return (more && !limit && !(hist.status & kPanicBits)) ? MODE_STEP : MODE_DONE;
  }

  // pull the next member (State::new, state.rs:52-81)
  __device__ int fetch(bool* exhausted) {
    const unsigned long long i = atomicAdd(A.queue, 1ull);
    if (i >= A.n_traj) {
      *exhausted = true;
      return MODE_DONE;
    }
    const size_t n_traj = A.n_traj;
    hist.reset(size_t(i));
#pragma unroll
    for (int n = 0; n < Sys::NP; ++n) prm[n] = A.params[size_t(n) * n_traj + i];
#pragma unroll
    for (int n = 0; n < NA; ++n) aux[n] = 0.0;
    t_curr = t_prev = A.t0;
    {
      double y0[N], p0[N];
#pragma unroll
      for (int n = 0; n < N; ++n) y0[n] = A.y0[size_t(n) * n_traj + i];
      const double hk = Sys::NP >= 4 ? prm[Sys::NP >= 4 ? 3 : 0] : 0.0;
      init_condition_eval<0, N>(A.history_id, A.t0, y0, hk, hist.status, p0);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        p_curr[n] = p_prev[n] = p0[n];
        d_curr[n] = 0.0;
        k[0][n] = 0.0;
      }
      hist.push(A.t0, p0, false, false, [](int, int) { return 0.0; });
      t_slide = -dev_inf();
    }
    disco.head = 0;
    disco.len = 0;
    if (HAS_PROP) disco.close_gaps();
    if (HAS_PROP)
      for (int q = 0; q < A.n_disco; ++q) disco.push(A.disco_t[q], A.disco_order[q]);
    has_last_call = 0u;
    det_valid = det_valid_prev = 0u;
#pragma unroll
    for (int l = 0; l < NLOC; ++l) {
      last_call[l] = 0.0;
      det_cache[l] = 0.0;
      const bool on_times = l < A.n_loc && A.loc[l].filter == DFB_FILTER_ON_TIMES;
      every_i[l] = on_times ? 0 : (l < A.n_loc ? A.loc[l].filter_n - 1 : 0);
      sep_prev[l] = on_times ? 0.0 : -1.7976931348623157e308;
      trk_t[l] = -1.7976931348623157e308;
      trk_order[l] = 0x7fffffff;  // usize::MAX
      trk_index[l] = 0;
    }
    n_steps = n_events = n_evlog = n_samples = n_iters = on_step_calls = 0u;
    n_make = n_k0 = 0u;
    if (A.on_start_cb != DFB_CB_NONE) {  // events_on_start.eval_mut (solver.rs:289)
      RefMut m{&t_curr, p_curr, d_curr, t_prev, p_prev, &hist};
      Sys::callback(A.on_start_cb, m, prm, aux);
    }
    on_step();  // events_on_step at t_init (solver.rs:290)
    return next_mode();
  }

  __device__ void store() {
    const size_t n_traj = A.n_traj, gid = hist.gid;
    uint32_t st = hist.status;
    if (t_curr == dev_inf()) st |= DFB_TRAJ_STOPPED;
    bool finite = true;
#pragma unroll
    for (int n = 0; n < N; ++n) finite = finite && (fabs(p_curr[n]) < dev_inf());
    if (!finite) st |= DFB_TRAJ_NONFINITE;
    A.t_final[gid] = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + gid] = p_curr[n];
    if (A.aux_final) {
#pragma unroll
      for (int n = 0; n < Sys::NA; ++n) A.aux_final[size_t(n) * n_traj + gid] = aux[n];
    }
    A.steps[gid] = n_steps;
    A.events[gid] = n_events;
    A.rhs_evals[gid] = (unsigned long long)n_make * S + n_k0;
    A.status[gid] = st;
    if (A.locate_iters) A.locate_iters[gid] = n_iters;
    if (A.n_events) A.n_events[gid] = n_evlog;
    if (A.n_samples) A.n_samples[gid] = n_samples;
  }

  // ---- Solver::run (solver.rs:275-313) as a warp-friendly state machine ------------------
  __device__ void run() {
    int mode = MODE_DONE;
    bool have = false, exhausted = false;
    double ev_time = 0.0;
    int ev_index = 0;
    uint32_t fired_mask = 0u;
    int parked_iters = 0;
    step_limit = A.max_steps > 0 ? (A.max_steps >= 0xFFFFFFFFll ? 0xFFFFFFFFu : unsigned(A.max_steps)) : 0xFFFFFFFFu;

    while (true) {
      while (mode == MODE_DONE && !exhausted) {
        if (have) store();
        mode = fetch(&exhausted);
        have = !exhausted;
      }
      const unsigned stepping = __ballot_sync(0xffffffffu, mode == MODE_STEP || mode == MODE_EVENT_RESTEP);
      const unsigned parked = __ballot_sync(0xffffffffu, mode == MODE_LOCATE);
      if (stepping == 0u && parked == 0u) break;

      // (1) batched event location (HListLocateEarliest, loc.rs:821-835)
      const bool locate_round =
          parked != 0u && (stepping == 0u || __popc(parked) >= A.locate_batch || parked_iters >= A.locate_wait);
      if (locate_round) {
        parked_iters = 0;
        if (mode == MODE_LOCATE) {
          bool found = false;
          double best_t = 0.0;
          int best_i = 0;
#pragma unroll
          for (int l = 0; l < NLOC; ++l) {
            if (l < A.n_loc && ((fired_mask >> l) & 1u)) {
              const double tl = locate_base(A.loc[l], l);
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
            // undo_step (state.rs:146-150); k[0] still holds d_prev
            t_curr = t_prev;
#pragma unroll
            for (int n = 0; n < N; ++n) {
              p_curr[n] = p_prev[n];
              d_curr[n] = k[0][n];
            }
            ev_time = best_t;
            ev_index = best_i;
            mode = MODE_EVENT_RESTEP;
          } else {
            commit_step(true);
            ++n_steps;
            on_step();
            mode = next_mode();
          }
        }
      } else if (parked != 0u) {
        ++parked_iters;
      }

      // (2) make_step rounds for every lane that can advance
      if (!locate_round) {
#pragma unroll 1
        for (int it = 0; it < A.inner_steps; ++it) {
          if (it > 0) {
            if (__ballot_sync(0xffffffffu, mode == MODE_STEP || mode == MODE_EVENT_RESTEP) == 0u) break;
            if (parked != 0u) ++parked_iters;
          }
          if (mode == MODE_STEP || mode == MODE_EVENT_RESTEP) {
            const bool restep = mode == MODE_EVENT_RESTEP;
            const double h_req = restep ? (ev_time - t_curr) : fmin(A.h, A.t1 - t_curr);
            if (t_prev == t_curr) {  // state.rs:97-98: make_step evaluates k[0] itself
              Sys::rhs(view_curr(), prm, d_curr);
              ++n_k0;
            }
            make_step(h_req);
            if (hist.status & kPanicBits) {
              mode = MODE_DONE;  // the reference panics here
            } else if (restep) {
              // solver.rs:304-309: commit; zero step; callback of the located event; commit
              commit_step(true);
              t_prev = t_curr;
#pragma unroll
              for (int n = 0; n < N; ++n) p_prev[n] = p_curr[n];
#pragma unroll
              for (int l = 0; l < NLOC; ++l)
                if (l == ev_index) fire_callback(A.loc[l], l);
              ++n_events;
              if (A.event_capacity > 0) {
                if (n_evlog < uint32_t(A.event_capacity)) {
                  A.ev_t[size_t(n_evlog) * A.n_traj + hist.gid] = ev_time;
                  A.ev_idx[size_t(n_evlog) * A.n_traj + hist.gid] = ev_index;
                }
                ++n_evlog;
              }
              commit_step(false);  // zero-length record: never selected by a lookup
              det_valid = 0u;      // the callback may have changed the state the detectors saw
              ++n_steps;
              on_step();
              mode = next_mode();
            } else {
              fired_mask = 0u;
              det_valid_prev = det_valid;
              det_valid = 0u;
#pragma unroll
              for (int l = 0; l < NLOC; ++l)
                if (l < A.n_loc && detect_phase(A.loc[l], l)) fired_mask |= 1u << l;
              if (hist.status & kPanicBits) {
                mode = MODE_DONE;
              } else if (fired_mask != 0u) {
                mode = MODE_LOCATE;
              } else {
                commit_step(true);
                ++n_steps;
                on_step();
                mode = next_mode();
              }
            }
          }
        }
      }
    }
  }
};

template <class Sys, int S, int I, int NLOC, bool HAS_PROP, bool SIMPLE, int KINDS, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
    dde_event_kernel(const __grid_constant__ DdeEventArgs A) {
  using Stepper = DdeEventStepper<Sys, S, I, NLOC, HAS_PROP, SIMPLE, KINDS>;
  __shared__ double disco_t[HAS_PROP ? (DFB_DISCO_CAP + 2 * NLOC) * BLOCK : 1];
  __shared__ int32_t disco_o[HAS_PROP ? DFB_DISCO_CAP * BLOCK : 1];
  Stepper st(A);
  const unsigned warp = (blockIdx.x * BLOCK + threadIdx.x) / 32u;  // < A.ring_warps (host-sized grid)
  constexpr unsigned kWarpRec = unsigned(Stepper::Hist::F) * 32u;
  st.hist.A = &A;
  st.hist.slot_stride = A.ring_warps * kWarpRec;
  st.hist.mask = unsigned(A.ring_capacity - 1);
  st.hist.col = A.ring + (size_t(warp) * kWarpRec + (threadIdx.x & 31u));
  st.disco.t_col = disco_t + (HAS_PROP ? threadIdx.x : 0);
  st.disco.o_col = disco_o + (HAS_PROP ? threadIdx.x : 0);
  st.disco.stride = BLOCK;
  st.disco.head = st.disco.len = 0;
  st.run();
}

}  // namespace DFB_NS
