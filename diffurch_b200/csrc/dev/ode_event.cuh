// Hot path #3: fixed-step explicit RK over an ODE ensemble WITH located events (reference:
// Solver::run solver.rs:292-310; Detect/Locate, detection and location methods, Periodic,
// filters, LocCallback, HListLocateEarliest, DedupLocF — loc.rs:11-80, 127-143, 200-335,
// 465-700, 795-883, 970-1052; State::{make_step, commit_step, make_zero_step, undo_step}
// state.rs:94-150).  BASELINE configs[3]: bouncing ball, egg billiard.
//
// Same semantics as the general integrator (integrator.cuh) for the ODE + state-function /
// Periodic locator subset, but built like the hot-path kernels:
//
//  * the whole State and every locator's bookkeeping live in registers: the locator loops are
//    unrolled over a compile-time locator count, so nothing is indexed at run time (the general
//    kernel keeps that state in local memory: 39 % LSU-pipe utilisation on the bouncing ball);
//  * ONE make_step call site: lanes taking an ordinary step and lanes re-taking a step truncated
//    to an event time run the same instructions (the step length is per lane);
//  * warp-level batching of event location: a lane that detects an event parks; the warp votes
//    and runs the ~45-iteration bisections together once enough lanes are parked (or the first
//    one has waited long enough, or nobody can advance), so diverging lanes do not serialise the
//    warp on every step;
//  * FAST arithmetic only: the step's interpolant is pre-combined ONCE per located event into
//    polynomial coefficients  y(theta) = p_prev + sum_j c_j theta^j,  c_j = h sum_i k_i bi[i][j],
//    so a bisection iteration is an (I-1)-term Horner chain per component instead of the
//    reference's S x I sums (rk.rs:24-47), and the derivative interpolant is skipped for
//    detectors that never read `s.d`.  STRICT keeps the reference's dense_output and its
//    operation order: bit-identical event times;
//  * dynamic trajectory assignment (global work counter) as in ode_adaptive.cuh: members stop at
//    different times (egg billiard: after 200 reflections), finished lanes pull the next member.
#pragma once
#ifndef DFB_EV_EAGER_MAX_N
#define DFB_EV_EAGER_MAX_N 2
#endif
#ifndef DFB_EV_SPEC_PAIRED
#define DFB_EV_SPEC_PAIRED 0  // two bisection iterations per round (3 evaluations): measured slower, see locate_spec
#endif
#ifndef DFB_EV_UNROLL
#define DFB_EV_UNROLL 1
#endif
#include "common.cuh"
#include "integrator.cuh"
#include "ode_fixed.cuh"
#include "systems.cuh"

namespace DFB_NS {

// SIMPLE = every locator is what `Locator::zero / above_zero / below_zero` builds (loc.rs:916-918):
// a state-function locator with a sign-change detection, located by Bisection, unfiltered.  That
// is the shape of the shipped event examples; knowing it at compile time removes the Periodic,
// filter, boolean-detection and other-location branches from the per-step detection code (which
// is ~45 % of the instructions of the bouncing-ball run in the generic profile).
// LEAN = no Trace policy on this problem (known at launch): the per-step `on_step` test disappears.
// SPEC != kEvSpecRuntime (SIMPLE only): the locators themselves are known at compile time — one byte per
// locator: detector id (bits 0-3), detection (bits 4-5: DFB_DET_ZERO / ABOVE_ZERO / BELOW_ZERO), DedupLocF
// (bit 6) — and n_loc == NLOC.  The detector switch of the system folds to the one function asked for (the
// bisection of the bouncing ball then interpolates x only, not x and v), and the per-step detection carries
// no descriptor loads, no selects on the detection kind and no tests of n_loc / dedup.
constexpr unsigned kEvSpecRuntime = 0xFFFFFFFFu;
__host__ __device__ constexpr unsigned ev_spec_byte(int detector_id, int detection, bool dedup) {
  return unsigned(detector_id & 15) | (unsigned(detection & 3) << 4) | (dedup ? 64u : 0u);
}
template <class Sys, int S_, int I_, unsigned long long AMASK, unsigned BMASK, bool FAST, int NLOC, bool SIMPLE = false,
          bool LEAN = false, unsigned SPEC = kEvSpecRuntime>
struct OdeEventStepper : OdeFixedStepper<Sys, S_, AMASK, BMASK, false, 1> {
  static constexpr bool HAS_SPEC = SIMPLE && SPEC != kEvSpecRuntime;
  __device__ __forceinline__ static constexpr int spec_id(int l) { return int((SPEC >> (8 * l)) & 15u); }
  __device__ __forceinline__ static constexpr int spec_detection(int l) { return int((SPEC >> (8 * l + 4)) & 3u); }
  __device__ __forceinline__ static constexpr bool spec_dedup(int l) { return ((SPEC >> (8 * l + 6)) & 1u) != 0u; }
  using Base = OdeFixedStepper<Sys, S_, AMASK, BMASK, false, 1>;
  static constexpr int S = S_, I = I_;
  static_assert(!Sys::reads_d, "views of this kernel carry the fresh k[0] where the reference carries the stale d");
  static constexpr int N = Sys::N;
  static constexpr int NA = Sys::NA > 0 ? Sys::NA : 1;
  using Ref = StateRef<N, NoHist>;
  using RefMut = StateRefMut<N, NoHist>;
  struct Tab {  // the continuous-extension rows of the argument block, for dense_output<D, N>
    static constexpr int S = S_, I = I_;
    const double* v;
    __device__ __forceinline__ double bi(int i, int j) const { return v[i * DFB_MAX_INTERP + j]; }
    __device__ __forceinline__ bool stage(int) const { return true; }  // exact shape (see ParamTab)
    __device__ __forceinline__ bool term(int) const { return true; }
  };
  using Base::A;
  using Base::t_curr;
  using Base::t_prev;

  // slot 0 of the base's [TPT = 1] arrays
  __device__ __forceinline__ double* p_curr() { return Base::p_curr[0]; }
  __device__ __forceinline__ double* p_prev() { return Base::p_prev[0]; }
  __device__ __forceinline__ double* d_curr() { return Base::d_curr[0]; }
  __device__ __forceinline__ double* prm() { return Base::prm[0]; }
  __device__ __forceinline__ double (*k())[N] { return Base::k[0]; }  // k()[0] = d_prev after a step

  size_t gid;
  double aux[NA];
  // FAST: a lane that detects an event pre-combines the step's interpolant into `cf` and saves d_prev
  // AT ONCE, inside the (rare) fired branch.  The stage vectors k[][] are then dead at the end of every
  // loop iteration, and `cf` / `dprev` are untouched on the event-free path — so the loop back-edge
  // carries no copies of them.  (Round 1 pre-combined lazily in the locate round, which kept all of
  // k[][] live around the loop: ~45 register moves per step, 24 % of the executed instructions.)
  // STRICT evaluates the reference's dense_output on k[][] itself and keeps it live.
  // Narrow states only (N <= 2: the bouncing ball): for N = 4 the extra 2 (I-1) N registers cost more than
  // the moves they save (egg billiard 15.7 -> 16.7 ms), so wide states keep round 1's scheme — coefficients
  // overwrite k[1 .. I-1], pre-combined lazily in the locate round.
  static constexpr bool EAGER = FAST && Sys::N <= DFB_EV_EAGER_MAX_N;
  static constexpr int NCF = I_ > 1 ? I_ - 1 : 1;
  double cf[EAGER ? NCF : 1][EAGER ? N : 1];
  double dprev[EAGER ? N : 1];
  double last_call[NLOC], sep_prev[NLOC];
  int every_i[NLOC];
  uint32_t has_last_call, status;
  bool combined;  // k[1 .. I-1] of the step in flight already hold the interpolant coefficients
  unsigned n_steps, n_events, n_evlog, n_samples;
  unsigned n_make, n_k0;  // make_step calls and k[0] evaluations: rhs_evals = S n_make + n_k0
  unsigned step_limit;    // max_steps as a 32-bit bound (0xFFFFFFFF = none), set once per kernel
  unsigned long long on_step_calls;
  unsigned n_iters;  // location-loop iterations of this member (dfb_totals.locate_iters)

  __device__ explicit OdeEventStepper(const OdeFixedArgs& A_) : Base(A_) {}

  __device__ __forceinline__ Ref view_curr() {
    return Ref{t_curr, p_curr(), d_curr(), t_prev, p_prev(), &this->nohist};
  }
  // state_fn.rs:181-188.  The view's `d` is d_prev: k[0] after a step (STRICT), the copy a parked lane saved
  // (FAST).  No functor this kernel takes reads it (static_assert above), so only its type matters.
  __device__ __forceinline__ Ref view_prev() {
    return Ref{t_prev, p_prev(), EAGER ? dprev : k()[0], t_prev, p_prev(), &this->nohist};
  }

  // ---- interpolant of the step in flight (State::eval, state.rs:83-92) ------------------
  __device__ __forceinline__ void precombine() {
    if (!FAST) return;
    static_assert(!FAST || EAGER || I <= S, "the interpolant coefficients are stored in k[1 .. I-1]");
    // y(t) = p_prev + sum_j c_j (t - t_prev)^j with c_j = h sum_i k_i bi[i][j] / h^j: the division
    // theta = (t - t_prev) / h of rk.rs/state.rs:85 is paid once per event, not once per iteration
    const double t_step = t_curr - t_prev;
    const double inv = 1.0 / t_step;
    double scale = 1.0;  // h / h^j
    double c[NCF][N];
#pragma unroll
    for (int j = 1; j < I; ++j) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < S; ++i) acc = fma(k()[i][n], A.bi[i * DFB_MAX_INTERP + j], acc);
        c[j - 1][n] = acc * scale;
      }
      scale *= inv;
    }
#pragma unroll
    for (int j = 1; j < I; ++j)
#pragma unroll
      for (int n = 0; n < N; ++n) {
        if (EAGER) cf[EAGER ? j - 1 : 0][EAGER ? n : 0] = c[j - 1][n];
        else k()[j][n] = c[j - 1][n];
      }
    if (EAGER) {
#pragma unroll
      for (int n = 0; n < N; ++n) dprev[EAGER ? n : 0] = k()[0][n];
    }
  }
  __device__ __forceinline__ double cj(int j, int n) { return EAGER ? cf[EAGER ? j : 0][EAGER ? n : 0] : k()[j + 1][n]; }
  template <int D>
  __device__ __forceinline__ void eval_in_step(double t, double* out) {
    if (FAST && D == 0) {
      const double s = t - t_prev;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = cj(I - 2, n);
#pragma unroll
        for (int j = I - 3; j >= 0; --j) acc = fma(acc, s, cj(j, n));
        out[n] = fma(acc, s, p_prev()[n]);
      }
    } else {
      const double t_step = t_curr - t_prev;
      const double theta = (t - t_prev) / t_step;
      dense_output<D, N>(Tab{A.bi}, p_prev(), t_step, theta, k(), out);
    }
  }

  // EvalState of locator L's function: which = 0 eval_curr, 1 eval_prev, 2 eval_at(t)
  __device__ __forceinline__ double loc_fn(const dfb_locator& L, int which, double t) {
    if (!SIMPLE && L.detection == DFB_DET_STEP) return 1.0;
    if (which == 0) return Sys::detector(L.detector_id, view_curr(), prm());
    if (which == 1) return Sys::detector(L.detector_id, view_prev(), prm());
    double y[N], dy[N];
    eval_in_step<0>(t, y);
    if (Sys::reads_d) eval_in_step<1>(t, dy);  // state_fn.rs:190-201 builds both; only `d` readers need it
    return Sys::detector(L.detector_id, Ref{t, y, Sys::reads_d ? dy : y, t, y, &this->nohist}, prm());
  }

  // detection_method (loc.rs:127-143), Periodic::detect (:318-322)
  __device__ __forceinline__ bool detect_base(const dfb_locator& L) {
    if (SIMPLE) {
      const double c = loc_fn(L, 0, 0.0);
      const double p = loc_fn(L, 1, 0.0);
      const bool up = c > 0.0 && p <= 0.0;    // AboveZero (loc.rs:130-132)
      const bool down_z = c <= 0.0 && p > 0.0;  // the falling half of Zero (loc.rs:127-129)
      const bool down = c < 0.0 && p >= 0.0;  // BelowZero (loc.rs:133-135)
      return L.detection == DFB_DET_ZERO ? (up || down_z) : (L.detection == DFB_DET_ABOVE_ZERO ? up : down);
    }
    if (L.kind == DFB_LOC_PERIODIC) {
      const double prev = floor((t_prev - L.offset) / L.period);
      const double curr = floor((t_curr - L.offset) / L.period);
      return prev < curr;
    }
    const int det = L.detection;
    if (det == DFB_DET_STEP) return true;
    const bool needs_prev = det == DFB_DET_ZERO || det == DFB_DET_ABOVE_ZERO || det == DFB_DET_BELOW_ZERO ||
                            det == DFB_DET_SWITCH || det == DFB_DET_SWITCH_TRUE || det == DFB_DET_SWITCH_FALSE;
    const double c = loc_fn(L, 0, 0.0);
    const double p = needs_prev ? loc_fn(L, 1, 0.0) : 0.0;
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
  __device__ double locate_base(const dfb_locator& L) {
    if (SIMPLE) {  // Bisection (loc.rs:235-259)
      double lft = t_prev, rgt = t_curr;
      if (loc_fn(L, 0, 0.0) < 0.0) {
        lft = t_curr;
        rgt = t_prev;
      }
      double w = fabs(rgt - lft);
      double w_prev = 2.0 * w;
      double m = 0.5 * (lft + rgt);
#pragma unroll 1
      while (w < w_prev) {
        w_prev = w;
        if (loc_fn(L, 2, m) < 0.0) lft = m;
        else rgt = m;
        m = 0.5 * (lft + rgt);
        w = fabs(rgt - lft);
        ++n_iters;
      }
      return fmax(lft, rgt);
    }
    if (L.kind == DFB_LOC_PERIODIC) return floor((t_curr - L.offset) / L.period) * L.period + L.offset;
    const int location = L.location;
    if (location == DFB_LOCATE_STEP_BEGIN) return t_prev;
    if (location == DFB_LOCATE_STEP_END) return t_curr;
    if (location == DFB_LOCATE_STEP_MIDDLE) return 0.5 * (t_curr - t_prev);  // sic (loc.rs:202-204)
    const bool is_bool = location == DFB_LOCATE_BISECTION_BOOL;
    const double fc = is_bool ? 0.0 : loc_fn(L, 0, 0.0);
    const double fp = (is_bool || location == DFB_LOCATE_LERP) ? loc_fn(L, 1, 0.0) : 0.0;
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
          fv[q] = loc_fn(L, 2, q == 0 ? lft : (q == 1 ? rgt : m));
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
      const double f = loc_fn(L, 2, m);
      const bool go_left = is_bool ? !(f != 0.0) : (f < 0.0);
      if (go_left) lft = m;
      else rgt = m;
      m = 0.5 * (lft + rgt);
      w = fabs(rgt - lft);
      ++n_iters;
    }
    return fmax(lft, rgt);
  }

  // Bisection (loc.rs:235-259) of locator LI of a SPEC build: the detector id is a constant.
  // DFB_EV_SPEC_PAIRED=1 (experiment, off): two iterations per round — next to f(m) the round evaluates f at
  // BOTH midpoints the following iteration could ask for, 0.5 (m + rgt) and 0.5 (lft + m), same expressions on
  // the same operands, so brackets, midpoints and iteration count are bit for bit those of the one-at-a-time
  // loop, and the three interpolations are independent instruction streams.  Measured: egg billiard 14.6 ->
  // 16.7 ms, bouncing ball 5.43 -> 5.60 ms — with all 32 lanes of a batched location round active and four
  // warps per scheduler the loop is bound by FP64 throughput, not by its dependent chain, and the third
  // evaluation is pure extra work (profiles/r2_event_kernel_tuning.txt).
  template <int LI>
  __device__ __forceinline__ double spec_fn(double t) {
    double y[N];
    eval_in_step<0>(t, y);
    return Sys::detector(spec_id(LI), Ref{t, y, y, t, y, &this->nohist}, prm());
  }
  template <int LI>
  __device__ double locate_spec() {
    double lft = t_prev, rgt = t_curr;
    if (Sys::detector(spec_id(LI), view_curr(), prm()) < 0.0) {
      lft = t_curr;
      rgt = t_prev;
    }
    double w = fabs(rgt - lft);
    double w_prev = 2.0 * w;
    double m = 0.5 * (lft + rgt);
#if DFB_EV_SPEC_PAIRED
#pragma unroll 1
    while (w < w_prev) {
      w_prev = w;
      const double m_l = 0.5 * (m + rgt);  // the next midpoint if lft = m
      const double m_r = 0.5 * (lft + m);  // ... if rgt = m
      const double f_m = spec_fn<LI>(m), f_l = spec_fn<LI>(m_l), f_r = spec_fn<LI>(m_r);
      const bool left = f_m < 0.0;
      lft = left ? m : lft;
      rgt = left ? rgt : m;
      const double m2 = left ? m_l : m_r;
      const double f2 = left ? f_l : f_r;
      w = fabs(rgt - lft);
      ++n_iters;
      m = m2;
      if (!(w < w_prev)) break;
      w_prev = w;
      if (f2 < 0.0) lft = m2;
      else rgt = m2;
      m = 0.5 * (lft + rgt);
      w = fabs(rgt - lft);
      ++n_iters;
    }
#else
#pragma unroll 1
    while (w < w_prev) {
      w_prev = w;
      if (spec_fn<LI>(m) < 0.0) lft = m;
      else rgt = m;
      m = 0.5 * (lft + rgt);
      w = fabs(rgt - lft);
      ++n_iters;
    }
#endif
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
      case DFB_FILTER_ON_TIMES: {  // loc.rs:665-688: every_i counts detections, sep_prev walks the list
        const int i = every_i[l];
        const int pos = int(sep_prev[l]);
        every_i[l] = i + 1;
        bool hit = false;
#pragma unroll
        for (int q = 0; q < DFB_MAX_FILTER_TIMES; ++q)  // unrolled: no run-time indexed array
          if (q == pos && q < L.filter_n && i == L.filter_times[q]) hit = true;
        if (hit) sep_prev[l] = double(pos + 1);
        return hit;
      }
    }
    return true;
  }

  // SIMPLE profile: every locator's detection (DedupLocF loc.rs:1002-1011 around a sign-change detection
  // loc.rs:127-135) as straight-line predicate arithmetic — no early outs, so no branches: a taken branch
  // costs this kernel ~15 issue cycles with 4 warps per scheduler, and round 1 had ~8 per step here
  __device__ __forceinline__ unsigned detect_all_simple() {
    unsigned fired = 0u;
#pragma unroll
    for (int l = 0; l < NLOC; ++l) {
      const dfb_locator& L = A.loc[l];
      const int id = HAS_SPEC ? spec_id(l) : L.detector_id;
      const int kind = HAS_SPEC ? spec_detection(l) : L.detection;
      const double c = Sys::detector(id, view_curr(), prm());
      const double p = Sys::detector(id, view_prev(), prm());
      const bool up = (c > 0.0) & (p <= 0.0);      // AboveZero (loc.rs:130-132)
      const bool down_z = (c <= 0.0) & (p > 0.0);  // the falling half of Zero (loc.rs:127-129)
      const bool down = (c < 0.0) & (p >= 0.0);    // BelowZero (loc.rs:133-135)
      const bool det = kind == DFB_DET_ZERO ? (up | down_z) : (kind == DFB_DET_ABOVE_ZERO ? up : down);
      const bool dedup = HAS_SPEC ? spec_dedup(l) : (L.dedup != 0);
      const bool suppressed = dedup & (((has_last_call >> l) & 1u) != 0u) & !(t_prev > last_call[l]);
      fired |= ((HAS_SPEC || l < A.n_loc) & det & !suppressed) ? (1u << l) : 0u;
    }
    return fired;
  }

  // DedupLocF (loc.rs:1002-1011) + filter wrappers (FilterAfterDetection :484-486, FilterLocated :568-576)
  __device__ __forceinline__ bool detect_phase(const dfb_locator& L, int l) {
    if (L.dedup && ((has_last_call >> l) & 1u) && !(t_prev > last_call[l])) return false;
    if (!detect_base(L)) return false;
    if (SIMPLE) return true;
    if (L.filter == DFB_FILTER_SEPARATED_BY || L.filter == DFB_FILTER_IN_INTERVAL) {
      if (!combined) precombine();  // FilterLocated::detect locates at once (rare: filtered locators only)
      combined = true;
      return filter_pass(L, l, locate_base(L));
    }
    if (L.filter == DFB_FILTER_EVERY || L.filter == DFB_FILTER_ON_TIMES) return filter_pass(L, l, t_curr);
    return true;
  }

  // output policy standing in for on_step closures (solver.rs:290, 309)
  __device__ __forceinline__ void on_step() {
    if (LEAN) return;
    if (A.trace_every > 0) {
      if (on_step_calls % (unsigned long long)A.trace_every == 0 && n_samples < uint32_t(A.trace_capacity)) {
        const size_t n_traj = A.n_traj;
        A.trace_t[size_t(n_samples) * n_traj + gid] = t_curr;
#pragma unroll
        for (int n = 0; n < N; ++n) A.trace_y[(size_t(n_samples) * N + n) * n_traj + gid] = p_curr()[n];
        ++n_samples;
      }
      ++on_step_calls;
    }
  }

  __device__ __forceinline__ int next_mode() {
    // branch-free: two return paths made the compiler duplicate the (predicated) register copies
    // that follow each call site
    const bool more = t_curr < A.t1;
    const bool limit = more && n_steps >= step_limit;
    status |= limit ? uint32_t(DFB_TRAJ_STEP_LIMIT) : 0u;
    return (more && !limit) ? MODE_STEP : MODE_DONE;
  }

  // pull the next member; returns MODE_DONE when the ensemble is exhausted
  __device__ int fetch(bool* exhausted) {
    const unsigned long long i = atomicAdd(A.queue, 1ull);
    if (i >= A.n_traj) {
      *exhausted = true;
      return MODE_DONE;
    }
    gid = size_t(i);
    const size_t n_traj = A.n_traj;
#pragma unroll
    for (int n = 0; n < Sys::NP; ++n) prm()[n] = A.params[size_t(n) * n_traj + i];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      p_curr()[n] = p_prev()[n] = A.y0[size_t(n) * n_traj + i];
      d_curr()[n] = 0.0;
      k()[0][n] = 0.0;
    }
#pragma unroll
    for (int n = 0; n < NA; ++n) aux[n] = 0.0;
    t_curr = t_prev = A.t0;
    has_last_call = 0u;
    status = 0u;
    combined = false;
#pragma unroll
    for (int l = 0; l < NLOC; ++l) {
      last_call[l] = 0.0;
      const bool on_times = l < A.n_loc && A.loc[l].filter == DFB_FILTER_ON_TIMES;
      every_i[l] = on_times ? 0 : (l < A.n_loc ? A.loc[l].filter_n - 1 : 0);
      sep_prev[l] = on_times ? 0.0 : -1.7976931348623157e308;  // T::min_value() / list position
    }
    n_steps = n_events = n_evlog = n_samples = n_iters = 0u;
    n_make = n_k0 = 0u;
    on_step_calls = 0ull;
    if (A.on_start_cb != DFB_CB_NONE) {  // events_on_start.eval_mut (solver.rs:289)
      RefMut m{&t_curr, p_curr(), d_curr(), t_prev, p_prev(), &this->nohist};
      Sys::callback(A.on_start_cb, m, prm(), aux);
    }
    on_step();  // events_on_step at t_init (solver.rs:290)
    return next_mode();
  }

  __device__ void store() {
    const size_t n_traj = A.n_traj;
    uint32_t st = status;
    if (t_curr == dev_inf()) st |= DFB_TRAJ_STOPPED;
    bool finite = true;
#pragma unroll
    for (int n = 0; n < N; ++n) finite = finite && (fabs(p_curr()[n]) < dev_inf());
    if (!finite) st |= DFB_TRAJ_NONFINITE;
    A.t_final[gid] = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + gid] = p_curr()[n];
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
    int parked_iters = 0;  // warp-uniform: step rounds since the first lane parked
    step_limit = A.max_steps > 0 ? (A.max_steps >= 0xFFFFFFFFll ? 0xFFFFFFFFu : unsigned(A.max_steps)) : 0xFFFFFFFFu;

    while (true) {
      // finished lanes hand in their member and pull the next one
      while (mode == MODE_DONE && !exhausted) {
        if (have) store();
        mode = fetch(&exhausted);
        have = !exhausted;
      }
      const unsigned stepping = __ballot_sync(0xffffffffu, mode == MODE_STEP || mode == MODE_EVENT_RESTEP);
      const unsigned parked = __ballot_sync(0xffffffffu, mode == MODE_LOCATE);
      if (stepping == 0u && parked == 0u) break;  // every lane is done and the ensemble is exhausted

      // (1) batched event location (HListLocateEarliest, loc.rs:821-835)
      const bool locate_round =
          parked != 0u && (stepping == 0u || __popc(parked) >= A.locate_batch || parked_iters >= A.locate_wait);
      if (locate_round) {
        parked_iters = 0;
        if (mode == MODE_LOCATE) {
          if (!EAGER) {  // EAGER lanes combined when they parked; a call here would keep k[][] live around the loop
            if (!combined) precombine();
            combined = true;
          }
          bool found = false;
          double best_t = 0.0;
          int best_i = 0;
          if (HAS_SPEC) {
            if (NLOC > 0 && (fired_mask & 1u)) {
              best_t = locate_spec<0>();
              found = true;
            }
            if (NLOC > 1 && (fired_mask & 2u)) {
              const double tl = locate_spec<(NLOC > 1 ? 1 : 0)>();
              if (!found || tl < best_t) {  // strictly earliest, first index wins ties
                found = true;
                best_t = tl;
                best_i = 1;
              }
            }
          } else {
#pragma unroll
            for (int l = 0; l < NLOC; ++l) {
              if (l < A.n_loc && ((fired_mask >> l) & 1u)) {
                const double tl = locate_base(A.loc[l]);
                if (!found || tl < best_t) {  // strictly earliest, first index wins ties
                  found = true;
                  best_t = tl;
                  best_i = l;
                }
              }
            }
          }
          if (found && best_t >= t_prev) {  // solver.rs:299-300
            // undo_step (state.rs:146-150)
            t_curr = t_prev;
#pragma unroll
            for (int n = 0; n < N; ++n) {
              p_curr()[n] = p_prev()[n];
              d_curr()[n] = EAGER ? dprev[EAGER ? n : 0] : k()[0][n];
            }
            ev_time = best_t;
            ev_index = best_i;
            mode = MODE_EVENT_RESTEP;
          } else {
            ++n_steps;  // commit_step
            on_step();
            mode = next_mode();
          }
        }
      } else if (parked != 0u) {
        ++parked_iters;
      }

      // (2) make_step rounds for every lane that can advance (not in a location round: the lanes
      // that just located join the next round's step together with everybody else).  Up to
      // A.inner_steps rounds run back to back before the warp votes again: lanes that park or
      // finish inside the burst idle for at most inner_steps - 1 rounds (they would wait for the
      // batch anyway), and the vote / batching / refill logic is paid once per burst.
      if (!locate_round) {
        constexpr int kBurstUnroll = DFB_EV_UNROLL;
#pragma unroll kBurstUnroll
        for (int it = 0; it < A.inner_steps; ++it) {
          if (it > 0) {
            if (__ballot_sync(0xffffffffu, mode == MODE_STEP || mode == MODE_EVENT_RESTEP) == 0u) break;
            if (parked != 0u) ++parked_iters;
          }
          if (mode == MODE_STEP || mode == MODE_EVENT_RESTEP) {
            const bool restep = mode == MODE_EVENT_RESTEP;
            const double h_req = restep ? (ev_time - t_curr) : fmin(A.h, A.t1 - t_curr);
            if (t_prev == t_curr) {  // state.rs:97-98: make_step evaluates k[0] itself
              Sys::rhs(view_curr(), prm(), d_curr());
              ++n_k0;
            }
            this->template step<false>(h_req);
            combined = false;
            ++n_make;
            if (restep) {
              // solver.rs:304-309: commit; zero step; callback of the located event; commit
              t_prev = t_curr;
#pragma unroll
              for (int n = 0; n < N; ++n) p_prev()[n] = p_curr()[n];
#pragma unroll
              for (int l = 0; l < NLOC; ++l) {
                if (l == ev_index) {
                  const dfb_locator& L = A.loc[l];
                  if (HAS_SPEC ? spec_dedup(l) : (L.dedup != 0)) {  // DedupLocF::eval_mut (loc.rs:1023-1026)
                    has_last_call |= 1u << l;
                    last_call[l] = t_curr;
                  }
                  if (L.callback_id != DFB_CB_NONE) {
                    RefMut m{&t_curr, p_curr(), d_curr(), t_prev, p_prev(), &this->nohist};
                    Sys::callback(L.callback_id, m, prm(), aux);
                  }
                }
              }
              ++n_events;
              if (A.event_capacity > 0) {
                if (n_evlog < uint32_t(A.event_capacity)) {
                  A.ev_t[size_t(n_evlog) * A.n_traj + gid] = ev_time;
                  A.ev_idx[size_t(n_evlog) * A.n_traj + gid] = ev_index;
                }
                ++n_evlog;
              }
              ++n_steps;
              on_step();
              mode = next_mode();
            } else {
              if (SIMPLE) {
                fired_mask = detect_all_simple();
              } else {
                fired_mask = 0u;
#pragma unroll
                for (int l = 0; l < NLOC; ++l)
                  if (l < A.n_loc && detect_phase(A.loc[l], l)) fired_mask |= 1u << l;
              }
              if (fired_mask != 0u) {
                if (EAGER && !combined) {  // park with the interpolant and d_prev in hand: k[][] dies here
                  precombine();
                  combined = true;
                }
                mode = MODE_LOCATE;
              } else {
                ++n_steps;  // commit_step
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

#ifndef DFB_EV_MINB
#define DFB_EV_MINB(N) ((N) <= 2 ? 4 : 3)
#endif
template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, bool FAST, int NLOC, int BLOCK,
          bool SIMPLE = false, bool LEAN = false, unsigned SPEC = kEvSpecRuntime>
__global__ void __launch_bounds__(BLOCK, DFB_EV_MINB(Sys::N))
    ode_event_kernel(const __grid_constant__ OdeFixedArgs A) {
  OdeEventStepper<Sys, S, I, AMASK, BMASK, FAST, NLOC, SIMPLE, LEAN, SPEC> st(A);
  st.run();
}

}  // namespace DFB_NS
