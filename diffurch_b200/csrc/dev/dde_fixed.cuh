// Hot path #2: fixed-step explicit RK over a constant-delay DDE/NDDE ensemble
// (reference: Solver::run solver.rs:292-310 + StateHistory state.rs:11-23,
// 122-139, 161-187 + dense_output rk.rs:17-52), FAST arithmetic.
//
// History store.  The reference keeps (t, y, k[S]) per step and re-evaluates the
// S x I continuous-extension sums on every lookup.  Here each committed step
// writes ONE record of pre-combined coefficients of the step's interpolant
//
//     y(t) = y_r + sum_{j=1}^{I-1} d_j (t - t_r)^j ,   d_j = sum_i k_i * bi[i][j] / h^(j-1)
//
// i.e. (y_r[N], d_1[N] .. d_{I-1}[N]) = 8 N I bytes (40 B for N = 1, I = 5) — the record's
// start time t_r is NOT stored: with a fixed step the time grid is t_{r+1} = t_r + h and the
// window below regenerates it bit for bit — into a per-trajectory ring in HBM laid out
// [slot][warp][field][lane]: every access of a warp is one fully coalesced 256-byte row, and
// one warp's record is a contiguous 8 N I x 32-byte block.  A lookup is then one
// subtraction and an (I-1)-term Horner chain; y'(t) for neutral equations comes
// from the same record.
//
// Streaming window.  With a constant delay tau the delayed times of one step,
// t + c_i h - tau, move monotonically through the history, so each thread keeps
// the two records that bracket them (r0, r1) in registers plus a third (r2)
// prefetched one step ahead; per step it stores one record and loads one:
// 2 x 8 N I bytes of HBM traffic per RK step (80 B for N = 1, I = 5), which is the kernel's bound
// (the ring of a 2^18-member ensemble is ~0.8 GB, far beyond L2).
#pragma once
#include "common.cuh"
#include "history_fn.cuh"
#include "systems.cuh"

namespace DFB_NS {

constexpr uint32_t kDdePanicBits =
    DFB_TRAJ_HISTORY_DELETED | DFB_TRAJ_HISTORY_NOT_READY | DFB_TRAJ_NO_DERIVATIVE;

struct DdeFixedArgs {
  double a[DFB_MAX_STAGES * DFB_MAX_STAGES];
  double b[DFB_MAX_STAGES];
  double c[DFB_MAX_STAGES];
  double ha[DFB_MAX_STAGES * DFB_MAX_STAGES];   // h * a_ij
  double hb[DFB_MAX_STAGES];                    // h * b_j
  double hc[DFB_MAX_STAGES];                    // c_i * h
  double bi[DFB_MAX_STAGES * DFB_MAX_INTERP];   // bi_ij
  double bih[DFB_MAX_STAGES * DFB_MAX_INTERP];  // bi_ij / h^(j-1)
  double t0, t1, h;
  double max_delay;       // Solver::max_delay: the retained history window (state.rs:126-133)
  double hc_min;          // min over stages i >= 1 of c_i * h (earliest delayed stage time of a step)
  double c_min;           // the same stage's c_i (for steps shorter than h)
  int32_t history_id;
  int32_t ring_capacity;  // power of two
  // full steps of length h that Solver::run takes before the closing partial step (solver.rs:292-293),
  // counted on the host by replaying t += h with the same IEEE additions; < 0 = not counted (the
  // kernel then evaluates the reference's own loop condition every step)
  long long n_full_steps;
  size_t n_traj;
  const double* y0;
  const double* params;
  double* t_final;
  double* y_final;
  unsigned long long* steps;
  unsigned long long* rhs_evals;
  uint32_t* status;
  double* ring;  // [cap][ceil(n/32)][N*I][32]
  // Trace output policy (on_step closures, solver.rs:290, 309); trace_every = 0: off
  int32_t trace_every, trace_capacity;
  double* trace_t;      // [cap][n]
  double* trace_y;      // [cap][N][n]
  uint32_t* n_samples;  // [n]
};

// ---- cp.async (LDGSTS) helpers: global -> shared without staging registers; completion is
// tracked by async groups, so an in-flight prefetch never blocks unrelated instructions.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_saddr(unsigned smem_addr, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}

#ifndef DFB_DDE_AHEAD
#define DFB_DDE_AHEAD 3  // 2 .. 5 (8 staging slots in the PRE build)
#endif
constexpr int kDdeSmemSlots = 4;  // records w .. w+2 of the window plus one being overwritten

// Horner evaluation of one staged record; p = this lane's column of the record block
// [field][32] in shared memory, s = t - t_r.
template <int D, int N, int I>
__device__ __forceinline__ void eval_staged(const double* p, double s, double* out) {
#pragma unroll
  for (int n = 0; n < N; ++n) {
    if (D == 0) {
      double acc = p[32 * (N * (I - 1) + n)];
#pragma unroll
      for (int j = I - 2; j >= 1; --j) acc = fma(acc, s, p[32 * (N * j + n)]);
      out[n] = fma(acc, s, p[32 * n]);
    } else {
      double acc = double(I - 1) * p[32 * (N * (I - 1) + n)];
#pragma unroll
      for (int j = I - 2; j >= 1; --j) acc = fma(acc, s, double(j) * p[32 * (N * j + n)]);
      out[n] = acc;
    }
  }
}

// The two-record window as the `history` of a StateRef.  STEADY = every lookup of the
// step is known to fall after t_init and inside the retained window (checked once per
// step by the stepper), so a lookup is a compare, a subtract and a Horner chain over
// coefficients read from the staged copy.
template <class Stepper, bool STEADY>
struct WindowHist {
  Stepper* st;
  template <int D>
  __device__ __forceinline__ void eval(double tt, double* out) {
    constexpr int N = Stepper::N, I = Stepper::I_;
    if (!STEADY) {
      if (tt <= st->A.t0) {  // state.rs:162-163: the history function (cold: state reloaded from HBM)
        double y0[N];
#pragma unroll
        for (int n = 0; n < N; ++n) y0[n] = st->A.y0[size_t(n) * st->A.n_traj + st->gid];
        const double hk = Stepper::SysT::NP >= 4 ? st->A.params[size_t(3) * st->A.n_traj + st->gid] : 0.0;
        init_condition_eval<D, N>(st->A.history_id, tt, y0, hk, st->status, out);
        return;
      }
      // the last committed time is the start of the step in flight
      if (tt >= st->t_prev) st->status |= DFB_TRAJ_HISTORY_NOT_READY;  // state.rs:172-177 (delay <= h)
      if (tt < st->t_oldest) st->status |= DFB_TRAJ_HISTORY_DELETED;   // state.rs:166-171 (max_delay < delay)
    }
    const bool second = tt >= st->tw1;
    const double* p = second ? st->p1 : st->p0;
    eval_staged<D, N, I>(p, tt - (second ? st->tw1 : st->tw0), out);
  }
};

// PRE: the delayed values of a whole steady step, evaluated BEFORE its stage chain.  With a constant delay
// the lookup times of a step, t_prev + c_i h - tau, do not depend on the state, so the S Horner chains are
// independent of each other and of the stages; taken out of the right-hand sides they run with ILP S, and the
// dependent chain of the step (stage sum -> rhs -> stage sum ...) is 2-3 FP64 latencies per stage instead of
// a staged-memory load + 4 dependent FMAs + 2-3.  That is what a LONE warp needs (a shard of the ensemble:
// 1-2 warps per SM sub-partition, nothing else hides the latency); a full device is issue-bound and does not
// care.  The functor still calls s.p_at(s.t - tau): the view's history hands back the value computed for that
// very time (same expression, same bits) and falls back to the window for any other.
template <class Stepper>
struct PreHist {
  Stepper* st;
  int i;  // stage slot (compile-time after unrolling)
  template <int D>
  __device__ __forceinline__ void eval(double tt, double* out) {
    constexpr int N = Stepper::N;
    if (tt == st->pre_t[i]) {
#pragma unroll
      for (int n = 0; n < N; ++n) out[n] = D == 0 ? st->pre_p[i][n] : st->pre_d[i][n];
    } else {
      WindowHist<Stepper, true> w{st};
      w.template eval<D>(tt, out);
    }
  }
};

template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, unsigned long long BIMASK, bool PRE = false>
struct DdeFixedStepper {
  using SysT = Sys;
  using Self = DdeFixedStepper;
  static constexpr int N = Sys::N;
  static constexpr int I_ = I;
  static constexpr int NP = Sys::NP > 0 ? Sys::NP : 1;
  static constexpr int F = N * I;          // doubles per record
  static constexpr int kWarpRec = F * 32;  // doubles of one warp's record block (F x 256 bytes)
  // staging slots: records w .. w+2 plus one being overwritten; the PRE build (lone warps: a shard of the
  // ensemble) prefetches one record deeper (w+3) so that a step never waits on HBM latency
  static constexpr int kSlots = PRE ? 8 : kDdeSmemSlots;
  static constexpr int kAhead = PRE ? DFB_DDE_AHEAD : 2;  // the lean loop fetches record w + kAhead

  const DdeFixedArgs& A;
  size_t gid;
  size_t g_thread;  // global thread index (= trajectory index for valid lanes)
  bool valid;
  int lane;
  double* sm;  // this warp's staging area [kSlots][F][32]
  double t_curr, t_prev;
  double p_curr[N], p_prev[N], d_curr[N];
  double k[S][N];
  double prm[NP];
  // window: records w (start tw0) and w+1 (start tw1); start times are regenerated as
  // t_{r+1} = t_r + h, which replays the shared time grid bit for bit
  double tw0, tw1;
  const double *p0, *p1;  // this lane's column of the two staged records
  double t_oldest;        // t_deque[0] of the reference after its prune
  double t_old1;          // its successor on the grid (t_deque[1])
  int w;
  uint32_t status;
  uint32_t n_samples = 0u, n_samples_alive = 0u;
  bool loaded, uniform;
  bool track_oldest;  // max_delay < tau: "deleted time range" panics (state.rs:166-171) are possible
  double tau_;        // this lane's delay (PRE)
  double pre_t[PRE ? S : 1], pre_p[PRE ? S : 1][N], pre_d[PRE ? S : 1][N];  // slot i-1 = stage i, slot S-1 = step end
  unsigned warp_base, slot_stride;  // ring addressing in doubles: [slot][warp][F][32] (host guarantees < 2^32)

  __device__ DdeFixedStepper(const DdeFixedArgs& A_, size_t g_, size_t gid_, bool valid_, double* sm_)
      : A(A_), gid(gid_), g_thread(g_), valid(valid_), lane(int(threadIdx.x & 31)), sm(sm_) {}

  template <bool STEADY>
  __device__ __forceinline__ void rhs(double* out) {
    WindowHist<Self, STEADY> h{this};
    Sys::rhs(StateRef<N, WindowHist<Self, STEADY>>{t_curr, p_curr, d_curr, t_prev, p_prev, &h}, prm, out);
  }
  __device__ __forceinline__ void rhs_pre(int slot, double* out) {
    PreHist<Self> h{this, slot};
    Sys::rhs(StateRef<N, PreHist<Self>>{t_curr, p_curr, d_curr, t_prev, p_prev, &h}, prm, out);
  }

  __device__ __forceinline__ static bool amask(int i, int j) { return (AMASK >> (i * 8 + j)) & 1ull; }

  template <bool FULL, bool STEADY>
  __device__ __forceinline__ void step(double t_step) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      k[0][n] = d_curr[n];
      p_prev[n] = p_curr[n];
    }
    t_prev = t_curr;
    constexpr bool kPre = PRE && FULL && STEADY;
    if (kPre) {
      WindowHist<Self, true> w{this};
#pragma unroll
      for (int i = 1; i <= S; ++i) {
        const double tt = (t_prev + (i < S ? A.hc[i] : t_step)) - tau_;
        pre_t[kPre ? i - 1 : 0] = tt;
        w.template eval<0>(tt, pre_p[kPre ? i - 1 : 0]);
        w.template eval<1>(tt, pre_d[kPre ? i - 1 : 0]);  // dead code unless the functor asks for y'(t - tau)
      }
    }
#pragma unroll
    for (int i = 1; i < S; ++i) {
      t_curr = t_prev + (FULL ? A.hc[i] : A.c[i] * t_step);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        if (FULL) {
          double acc = p_prev[n];
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (amask(i, j)) acc = fma(k[j][n], A.ha[i * DFB_MAX_STAGES + j], acc);
          p_curr[n] = acc;
        } else {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (amask(i, j)) acc = fma(k[j][n], A.a[i * DFB_MAX_STAGES + j], acc);
          p_curr[n] = fma(acc, t_step, p_prev[n]);
        }
      }
      if (kPre) rhs_pre(i - 1, k[i]);
      else rhs<STEADY>(k[i]);
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
      if (FULL) {
        double acc = p_prev[n];
#pragma unroll
        for (int j = 0; j < S; ++j)
          if (BMASK & (1u << j)) acc = fma(k[j][n], A.hb[j], acc);
        p_curr[n] = acc;
      } else {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < S; ++j)
          if (BMASK & (1u << j)) acc = fma(k[j][n], A.b[j], acc);
        p_curr[n] = fma(acc, t_step, p_prev[n]);
      }
    }
    t_curr = t_prev + t_step;
    if (kPre) rhs_pre(S - 1, d_curr);
    else rhs<STEADY>(d_curr);
  }

  // ---- ring I/O: [slot][warp][field][lane] -----------------------------------
  // One warp's record is a contiguous F x 256-byte block: every field access is one
  // fully coalesced row, field offsets are compile-time immediates, and the block moves
  // to shared memory as F x 16 16-byte cp.async chunks (F/2 per lane, + half a warp if F is odd).
  __device__ __forceinline__ double* warp_block(int rec) const {
    const unsigned slot = unsigned(rec) & unsigned(A.ring_capacity - 1);
    return A.ring + (slot * slot_stride + warp_base);
  }
  __device__ __forceinline__ double* smem_block(int rec) const {
    return sm + (unsigned(rec) & unsigned(kSlots - 1)) * unsigned(kWarpRec);
  }
  __device__ __forceinline__ void issue_fetch(int rec) {
    const double* g = warp_block(rec);
    double* d = smem_block(rec);
#pragma unroll
    for (int q = 0; q < F / 2; ++q) cp_async16(d + 2 * (lane + 32 * q), g + 2 * (lane + 32 * q));
    if ((F & 1) && lane < 16) cp_async16(d + 2 * (lane + 32 * (F / 2)), g + 2 * (lane + 32 * (F / 2)));
  }
  // issue_fetch with the lane's shared-memory address and ring column computed once (lean steady loop)
  __device__ __forceinline__ void issue_fetch_lean(int rec, unsigned sm_lane, const double* ring_lane) {
    const double* g = ring_lane + (unsigned(rec) & unsigned(A.ring_capacity - 1)) * slot_stride;
    const unsigned d = sm_lane + (unsigned(rec) & unsigned(kSlots - 1)) * unsigned(kWarpRec * sizeof(double));
#pragma unroll
    for (int q = 0; q < F / 2; ++q) cp_async16_saddr(d + 512u * q, g + 64 * q);
    if ((F & 1) && lane < 16) cp_async16_saddr(d + 512u * (F / 2), g + 64 * (F / 2));
  }
  // per-lane copy of one record (lanes of the warp disagree on the window position)
  __device__ __forceinline__ void copy_own_column(int rec) {
    const double* g = warp_block(rec) + lane;
    double* d = smem_block(rec) + lane;
#pragma unroll
    for (int f = 0; f < F; ++f) d[32 * f] = __ldcs(g + 32 * f);
  }

  // commit_step (state.rs:122-139): pre-combine the interpolant of the step just
  // taken and stream it to the ring.  The ring overwrites its oldest slot, which is
  // the prune rule for a window of max_delay.
  template <bool FULL>
  __device__ __forceinline__ void commit(int s, double t_step) {
    double inv = 1.0;  // 1 / h^(j-1) for partial steps
    double* p = warp_block(s) + lane;
#pragma unroll
    for (int n = 0; n < N; ++n) __stcs(p + 32 * n, p_prev[n]);
#pragma unroll
    for (int j = 1; j < I; ++j) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = 0.0;
        bool first = true;
#pragma unroll
        for (int i = 0; i < S; ++i) {
          if (!((BIMASK >> (i * 8 + j)) & 1ull)) continue;  // structural zero of bi
          const double cf = FULL ? A.bih[i * DFB_MAX_INTERP + j] : A.bi[i * DFB_MAX_INTERP + j];
          acc = first ? k[i][n] * cf : fma(k[i][n], cf, acc);
          first = false;
        }
        __stcs(p + 32 * (N * j + n), FULL ? acc : acc * inv);
      }
      if (!FULL) inv = inv / t_step;
    }
  }

  // After committing record s: replay the reference's prune, slide the window so that
  // records (w, w+1) bracket the delayed times of the NEXT step, and prefetch one ahead.
  // Invariant of the staged path: after refill(s) every record <= min(w + 2, s) has been
  // issued to shared memory, the newest of them possibly still in flight.
  __device__ __forceinline__ void refill(int s, double tau, double h) {
    // prune rule of commit_step (state.rs:126-133) replayed on the shared time grid:
    // pop while t_deque[1] < t_prev - max_delay
    // (needed only when the retained window may be shorter than the delay: with max_delay >= tau
    // no lookup can fall before the oldest retained record, see run())
    if (track_oldest) {
      const double t_tail = t_prev - A.max_delay;
      while (t_old1 < t_tail && t_old1 < t_prev) {
        t_oldest = t_old1;
        t_old1 = t_old1 + h;
      }
    }
    int issued_before;  // highest record issued by earlier refills
    if (!loaded) {
      if (!(t_curr + h - tau > A.t0)) return;  // every lookup of the next step is in the history function
      loaded = true;
      w = 0;
      tw0 = A.t0;
      tw1 = A.t0 + h;
      issued_before = -1;
    } else {
      issued_before = (w + 2 < s - 1) ? w + 2 : s - 1;
    }
    // Earliest lookup of the NEXT step: stage 0 reuses d_curr (no lookup), so it is the stage with
    // the smallest c_i > 0, evaluated with the very expression that stage will use.  Sliding on it
    // (not on t_curr - tau) puts every lookup of the step in [T_w, T_w+2): they span less than
    // (1 - c_min) h.  Sliding on t_curr - tau let the end-of-step lookup reach T_w+2 by one ulp when
    // tau / h is an integer, and the window then extrapolated record w+1 where the reference reads
    // record w+2 — invisible for a smooth interpolant, an O(h) error for y'(t - tau) of a linear one.
    const double next_len = fmin(h, A.t1 - t_curr);
    const double tt_min = (t_curr + (next_len == h ? A.hc_min : A.c_min * next_len)) - tau;
    while (tw1 <= tt_min) {              // record w+1 starts at or before it: slide
      ++w;
      tw0 = tw1;
      tw1 = tw1 + h;
    }
    if (uniform) {
      const int want = (w + 2 < s) ? w + 2 : s;
      // a 16-byte chunk holds the columns of two lanes: when the record just committed is
      // fetched straight back (short delays), the other lanes' stores must be ordered first
      if (want >= s) __syncwarp();
#pragma unroll 1
      for (int r = issued_before + 1; r <= want; ++r) issue_fetch(r);
      cp_async_commit();
      const int need = (w + 1 < s) ? w + 1 : s;        // newest record the next step reads
      if (need <= issued_before) cp_async_wait<1>();   // issued by an earlier refill: one step of slack
      else cp_async_wait<0>();                         // start-up / short delays: issued just now
      __syncwarp();
    } else {
      copy_own_column(w);
      if (w + 1 <= s) copy_own_column(w + 1);
    }
    p0 = smem_block(w) + lane;
    p1 = smem_block(w + 1) + lane;
    // If record w+1 is the step about to be taken (short delays) its start time tw1 equals the
    // next step's t_prev, so only lookups the reference rejects as "not yet computed" select it.
  }

  // refill for the steady regime of a warp-uniform delay (every record the window fetches is at least three
  // steps older than the one just committed, no prune replay): the same window rule as refill() without its
  // start-up, short-delay and per-lane branches.  Anything else goes through refill().
  __device__ __forceinline__ void refill_steady(int s, double tau, double h) {
    if (!(uniform && !track_oldest && loaded && w + 4 < s)) {
      refill(s, tau, h);
      return;
    }
    const double next_len = fmin(h, A.t1 - t_curr);
    const double tt_min = (t_curr + (next_len == h ? A.hc_min : A.c_min * next_len)) - tau;
    int r = w + 3;  // records <= w + 2 were issued by earlier refills
    while (tw1 <= tt_min) {
      ++w;
      tw0 = tw1;
      tw1 = tw1 + h;
    }
    const bool jumped = w + 2 > r;  // moved by more than one record (rounding of the grid, rare)
#pragma unroll 1
    for (; r <= w + 2 && r <= s; ++r) issue_fetch(r);
    cp_async_commit();
    if (jumped) cp_async_wait<0>();  // its records may have been issued only now
    else cp_async_wait<1>();  // record w + 1 was issued one refill ago at the latest
    __syncwarp();
    p0 = smem_block(w) + lane;
    p1 = smem_block(w + 1) + lane;
  }

  // every trace_every-th on_step invocation stores (t, p); the grid is shared, the test is uniform
  __device__ __forceinline__ void on_step(unsigned calls) {
    if (A.trace_every > 0 && calls % unsigned(A.trace_every) == 0u && n_samples < uint32_t(A.trace_capacity)) {
      if (valid) {
        A.trace_t[size_t(n_samples) * A.n_traj + gid] = t_curr;
#pragma unroll
        for (int n = 0; n < N; ++n) A.trace_y[(size_t(n_samples) * N + n) * A.n_traj + gid] = p_curr[n];
      }
      ++n_samples;
      // the reference stops at a panic: what a panicked member stores from then on is not counted
      if (!(status & kDdePanicBits)) n_samples_alive = n_samples;
    }
  }

  __device__ void run() {
    const size_t n_traj = A.n_traj;
#pragma unroll
    for (int n = 0; n < Sys::NP; ++n) prm[n] = A.params[size_t(n) * n_traj + gid];
    const double tau = Sys::const_delay(prm);
    tau_ = tau;
    // the staged (warp-cooperative) path needs one window position per warp
    uniform = __all_sync(0xffffffffu, __double_as_longlong(tau) ==
                                          __shfl_sync(0xffffffffu, __double_as_longlong(tau), 0));
    // The reference's prune leaves a front record older than t_prev - max_delay - h; every lookup of
    // this system is at t - tau >= t_prev - tau.  With max_delay >= tau the front can never be
    // newer than a lookup, so the replay of the prune (and its per-step check) is skipped.
    track_oldest = !(A.max_delay >= tau);
    const size_t n_warps = (n_traj + 31) / 32;
    slot_stride = unsigned(n_warps * size_t(kWarpRec));
    warp_base = unsigned((g_thread / 32) * size_t(kWarpRec));
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    t_oldest = A.t0;
    t_old1 = A.t0 + A.h;
    tw0 = tw1 = inf;
    p0 = p1 = sm + lane;
    status = 0u;
    loaded = false;
    w = 0;
    // State::new (state.rs:52-81)
    t_curr = t_prev = A.t0;
    {
      double y0[N], p0v[N];
#pragma unroll
      for (int n = 0; n < N; ++n) y0[n] = A.y0[size_t(n) * n_traj + gid];
      const double hk = Sys::NP >= 4 ? prm[Sys::NP >= 4 ? 3 : 0] : 0.0;
      init_condition_eval<0, N>(A.history_id, A.t0, y0, hk, status, p0v);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        p_curr[n] = p_prev[n] = p0v[n];
        d_curr[n] = 0.0;
      }
    }
    int s = 0;  // index of the step being taken = ring record it will write
    n_samples = n_samples_alive = 0u;
    on_step(0u);  // events_on_step at t_init (solver.rs:290)
    // A trajectory that hits one of the reference's panics keeps stepping (its warp must
    // stay converged for the cooperative staging); its status word records the panic.
    if (t_curr < A.t1) {
      rhs<false>(d_curr);  // k[0] of the first step (state.rs:97-98)
      const double h = A.h;
      // Phase A: some lookup of the step may still reach the history function, or the
      // delay does not exceed the step: every lookup carries the reference's checks.
      const bool counted = A.n_full_steps >= 0;
      const int n_full = int(A.n_full_steps);
#pragma unroll 1
      while ((counted ? s < n_full : (t_curr < A.t1 && (A.t1 - t_curr) >= h)) &&
             !(loaded && (t_curr + A.hc_min) - tau > A.t0 && (t_curr + A.hc_min) - tau >= tw0 &&
               (t_curr + h) - tau < t_curr)) {
        step<true, false>(h);
        commit<true>(s, h);
        refill(s, tau, h);
        ++s;
        on_step(unsigned(s));
      }
      // Phase B, lean form: while the step AFTER the one being taken is a counted full step as well, the
      // delay is warp-uniform, no prune replay and no trace are asked for, the per-step bookkeeping of
      // refill_steady() (same window rule, same invariants) shrinks to: two additions for the earliest
      // lookup time of the next step, one predicated slide, one record fetch, the async-group wait.  A
      // shard of the ensemble (1-2 warps per SM sub-partition) spent half of its time in the branches
      // and dependent selects of the general form (profiles/r2_dde_small_shard_lean_loop.txt).
      if (counted && uniform && !track_oldest && loaded && A.trace_every <= 0) {
        const unsigned sm_lane = (unsigned)__cvta_generic_to_shared(sm) + 16u * unsigned(lane);
        const double* ring_lane = A.ring + warp_base + 2 * lane;
        // records <= w + 2 are issued (invariant of refill); the deeper prefetch starts with the ones up to w + kAhead
        if (s + 1 < n_full && w + kAhead + 2 < s) {
          for (int r = w + 3; r <= w + kAhead; ++r) issue_fetch_lean(r, sm_lane, ring_lane);
          cp_async_commit();
        }
#pragma unroll 1
        while (s + 1 < n_full && w + kAhead + 2 < s) {
          step<true, true>(h);
          commit<true>(s, h);
          const double tt_min = (t_curr + A.hc_min) - tau;
          const int w_was = w;
          if (tw1 <= tt_min) {
            ++w;
            tw0 = tw1;
            tw1 = tw1 + h;
          }
          if (tw1 <= tt_min) {  // more than one record per step: rounding of the grid, rare
            do {
              ++w;
              tw0 = tw1;
              tw1 = tw1 + h;
            } while (tw1 <= tt_min);
            for (int r = w_was + kAhead + 1; r < w + kAhead; ++r) issue_fetch_lean(r, sm_lane, ring_lane);
            cp_async_commit();
            cp_async_wait<0>();  // the window jumped: its records may have been issued only now
          }
          if (w != w_was) issue_fetch_lean(w + kAhead, sm_lane, ring_lane);  // committed long ago (w + kAhead + 2 < s)
          cp_async_commit();
          cp_async_wait<kAhead - 1>();  // record w + 1 was issued kAhead - 1 steps ago at the latest
          __syncwarp();
          p0 = smem_block(w) + lane;
          p1 = smem_block(w + 1) + lane;
          ++s;
        }
      }
      // Phase B (steady): full steps whose lookups all lie after t_init and before the
      // last committed time.  What the reference checks on every lookup is checked once
      // per step: the earliest delayed stage time against the oldest retained record
      // (state.rs:166-171).
#pragma unroll 1
      while (counted ? s < n_full : (t_curr < A.t1 && (A.t1 - t_curr) >= h)) {
        if (track_oldest && (t_curr + A.hc_min) - tau < t_oldest) status |= DFB_TRAJ_HISTORY_DELETED;
        step<true, true>(h);
        commit<true>(s, h);
        refill_steady(s, tau, h);
        ++s;
        on_step(unsigned(s));
      }
      // closing partial step(s): checked path
#pragma unroll 1
      while (t_curr < A.t1) {
        const double h_req = fmin(h, A.t1 - t_curr);
        step<false, false>(h_req);
        commit<false>(s, h_req);
        refill(s, tau, h);
        ++s;
        on_step(unsigned(s));
      }
    }
    cp_async_wait<0>();
    if (!valid) return;
    bool finite = true;
#pragma unroll
    for (int n = 0; n < N; ++n) finite = finite && (fabs(p_curr[n]) < inf);
    if (!finite && status == 0u) status |= DFB_TRAJ_NONFINITE;
    A.t_final[gid] = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + gid] = p_curr[n];
    A.steps[gid] = (unsigned long long)s;
    A.rhs_evals[gid] = s > 0 ? (unsigned long long)s * S + 1 : 0;
    A.status[gid] = status;
    if (A.trace_every > 0 && A.n_samples) A.n_samples[gid] = n_samples_alive;
  }
};

template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, unsigned long long BIMASK,
          int BLOCK, int MIN_BLOCKS, bool PRE = false>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
    dde_fixed_kernel(const __grid_constant__ DdeFixedArgs A) {
  using Stepper = DdeFixedStepper<Sys, S, I, AMASK, BMASK, BIMASK, PRE>;
  // (blockDim.x / 32) * Stepper::kSlots * kWarpRec doubles, sized by the launcher: BLOCK is the launch
  // bound, the block itself is chosen from the ensemble size (host/launch_geometry.hpp)
  extern __shared__ __align__(16) double smem[];
  const size_t g = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  // every lane of a warp stays in the kernel (warp-cooperative staging); lanes past the
  // end replay the last trajectory and store nothing.  The ring is allocated for whole warps.
  if ((g & ~size_t(31)) >= A.n_traj) return;  // whole warp past the end (warp-uniform exit)
  const bool valid = g < A.n_traj;
  const size_t gid = valid ? g : A.n_traj - 1;
  Stepper st(A, g, gid, valid, smem + (threadIdx.x / 32) * Stepper::kSlots * Stepper::kWarpRec);
  st.run();
}

}  // namespace DFB_NS
