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
// i.e. (t_r, y_r[N], d_1[N] .. d_{I-1}[N]) = 8 (1 + N I) bytes (48 B for N = 1, I = 5),
// into a per-trajectory ring in HBM laid out [slot][field][trajectory]: every
// access of a warp is one fully coalesced 256-byte row.  A lookup is then one
// subtraction and an (I-1)-term Horner chain; y'(t) for neutral equations comes
// from the same record.
//
// Streaming window.  With a constant delay tau the delayed times of one step,
// t + c_i h - tau, move monotonically through the history, so each thread keeps
// the two records that bracket them (r0, r1) in registers plus a third (r2)
// prefetched one step ahead; per step it stores one record and loads one:
// 2 x 8(1 + N I) bytes of HBM traffic per RK step, which is the kernel's bound
// (the ring of a 2^18-member ensemble is ~0.8 GB, far beyond L2).
#pragma once
#include "common.cuh"
#include "history_fn.cuh"
#include "systems.cuh"

namespace DFB_NS {

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
  int32_t history_id;
  int32_t ring_capacity;  // power of two
  size_t n_traj;
  const double* y0;
  const double* params;
  double* t_final;
  double* y_final;
  unsigned long long* steps;
  unsigned long long* rhs_evals;
  uint32_t* status;
  double* ring;  // [cap][1 + N*I][n]
};

template <int N, int I>
struct CoeffRec {
  double t;
  double y[N];
  double d[I > 1 ? I - 1 : 1][N];
};

// The two-record window as the `history` of a StateRef.
template <int N, int I>
struct WindowHist {
  CoeffRec<N, I> r0, r1;
  double t_init, t_commit;  // t_commit = latest committed time (lookups at or after it are "not yet computed")
  double t_oldest;          // t_deque[0] of the reference after its prune (lookups before it are "deleted")
  double y0[N];
  double hist_k;
  int history_id;
  uint32_t status;

  template <int D>
  __device__ __forceinline__ void eval(double tt, double* out) {
    if (tt <= t_init) {  // state.rs:162-163
      init_condition_eval<D, N>(history_id, tt, y0, hist_k, status, out);
      return;
    }
    if (tt >= t_commit) status |= DFB_TRAJ_HISTORY_NOT_READY;  // state.rs:172-177 (delay <= h)
    if (tt < t_oldest) status |= DFB_TRAJ_HISTORY_DELETED;     // state.rs:166-171 (max_delay < delay)
    const bool second = tt >= r1.t;
    const double s = tt - (second ? r1.t : r0.t);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      if (D == 0) {
        double acc = second ? r1.d[I - 2][n] : r0.d[I - 2][n];
#pragma unroll
        for (int j = I - 3; j >= 0; --j) acc = fma(acc, s, second ? r1.d[j][n] : r0.d[j][n]);
        out[n] = fma(acc, s, second ? r1.y[n] : r0.y[n]);
      } else {
        double acc = double(I - 1) * (second ? r1.d[I - 2][n] : r0.d[I - 2][n]);
#pragma unroll
        for (int j = I - 3; j >= 0; --j)
          acc = fma(acc, s, double(j + 1) * (second ? r1.d[j][n] : r0.d[j][n]));
        out[n] = acc;
      }
    }
  }
};

template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, unsigned long long BIMASK>
struct DdeFixedStepper {
  static constexpr int N = Sys::N;
  static constexpr int NP = Sys::NP > 0 ? Sys::NP : 1;
  static constexpr int F = 1 + N * I;  // doubles per record
  using Hist = WindowHist<N, I>;
  using Ref = StateRef<N, Hist>;
  using Rec = CoeffRec<N, I>;

  const DdeFixedArgs& A;
  size_t gid;
  double t_curr, t_prev;
  double p_curr[N], p_prev[N], d_curr[N];
  double k[S][N];
  double prm[NP];
  Hist hist;
  Rec r2;          // prefetched
  double t_old1;   // t_deque[1] of the reference (successor of hist.t_oldest on the shared time grid)
  long long w;     // interval index held in hist.r0
  bool loaded;

  __device__ DdeFixedStepper(const DdeFixedArgs& A_, size_t gid_) : A(A_), gid(gid_) {}

  __device__ __forceinline__ void rhs(double* out) {
    Sys::rhs(Ref{t_curr, p_curr, d_curr, t_prev, p_prev, &hist}, prm, out);
  }

  __device__ __forceinline__ static bool amask(int i, int j) { return (AMASK >> (i * 8 + j)) & 1ull; }

  template <bool FULL>
  __device__ __forceinline__ void step(double t_step) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      k[0][n] = d_curr[n];
      p_prev[n] = p_curr[n];
    }
    t_prev = t_curr;
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
      rhs(k[i]);
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
    rhs(d_curr);
  }

  // ---- ring I/O: [slot][field][trajectory] ---------------------------------
  __device__ __forceinline__ size_t field_addr(long long rec, int field) const {
    const size_t slot = size_t(rec) & size_t(A.ring_capacity - 1);
    return (slot * F + size_t(field)) * A.n_traj + gid;
  }
  __device__ __forceinline__ Rec load_rec(long long rec, long long newest) const {
    Rec r;
    if (rec > newest) {  // not written yet: never selected by a lookup
      r.t = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        r.y[n] = 0.0;
#pragma unroll
        for (int j = 0; j < I - 1; ++j) r.d[j][n] = 0.0;
      }
      return r;
    }
    r.t = __ldcs(A.ring + field_addr(rec, 0));
#pragma unroll
    for (int n = 0; n < N; ++n) r.y[n] = __ldcs(A.ring + field_addr(rec, 1 + n));
#pragma unroll
    for (int j = 0; j < I - 1; ++j)
#pragma unroll
      for (int n = 0; n < N; ++n) r.d[j][n] = __ldcs(A.ring + field_addr(rec, 1 + N * (j + 1) + n));
    return r;
  }

  // commit_step (state.rs:122-139): pre-combine the interpolant of the step just
  // taken and stream it to the ring.  The ring overwrites its oldest slot, which is
  // the prune rule for a window of max_delay.
  template <bool FULL>
  __device__ __forceinline__ void commit(long long s, double t_step) {
    double inv = 1.0;  // 1 / h^(j-1) for partial steps
    __stcs(A.ring + field_addr(s, 0), t_prev);
#pragma unroll
    for (int n = 0; n < N; ++n) __stcs(A.ring + field_addr(s, 1 + n), p_prev[n]);
#pragma unroll
    for (int j = 1; j < I; ++j) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < S; ++i) {
          if (!((BIMASK >> (i * 8 + j)) & 1ull)) continue;  // structural zero of bi
          const double cf = FULL ? A.bih[i * DFB_MAX_INTERP + j] : A.bi[i * DFB_MAX_INTERP + j];
          acc = fma(k[i][n], cf, acc);
        }
        __stcs(A.ring + field_addr(s, 1 + N * j + n), FULL ? acc : acc * inv);
      }
      if (!FULL) inv = inv / t_step;
    }
  }

  // keep (r0, r1) bracketing the delayed times of the NEXT step, r2 one record ahead
  __device__ __forceinline__ void refill(long long s, double tau, double h_next) {
    hist.t_commit = t_curr;
    // prune rule of commit_step (state.rs:126-133) replayed on the shared time grid:
    // pop while t_deque[1] < t_prev - max_delay.  t_{r+1} = t_r + h reproduces the grid bit for bit.
    const double t_tail = t_prev - A.max_delay;
    while (t_old1 < t_tail && t_old1 < t_prev) {
      hist.t_oldest = t_old1;
      t_old1 = t_old1 + h_next;
    }
    if (!loaded) {
      if (t_curr + h_next - tau > A.t0) {
        w = 0;
        hist.r0 = load_rec(0, s);
        hist.r1 = load_rec(1, s);
        r2 = load_rec(2, s);
        loaded = true;
      }
      return;
    }
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (hist.r1.t == inf && w + 1 <= s) hist.r1 = load_rec(w + 1, s);
    if (r2.t == inf && w + 2 <= s) r2 = load_rec(w + 2, s);
    const double tt_min = t_curr - tau;  // earliest delayed time of the next step (c_0 = 0)
    while (hist.r1.t <= tt_min) {
      hist.r0 = hist.r1;
      hist.r1 = r2;
      ++w;
      r2 = load_rec(w + 2, s);
    }
  }

  __device__ void run() {
    const size_t n_traj = A.n_traj;
#pragma unroll
    for (int n = 0; n < Sys::NP; ++n) prm[n] = A.params[size_t(n) * n_traj + gid];
    const double tau = Sys::const_delay(prm);
    hist.t_init = A.t0;
    hist.t_commit = A.t0;
    hist.t_oldest = A.t0;
    t_old1 = A.t0 + A.h;
    hist.history_id = A.history_id;
    hist.hist_k = Sys::NP >= 4 ? prm[Sys::NP >= 4 ? 3 : 0] : 0.0;
    hist.status = 0u;
#pragma unroll
    for (int n = 0; n < N; ++n) hist.y0[n] = A.y0[size_t(n) * n_traj + gid];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    hist.r0.t = hist.r1.t = r2.t = inf;
    loaded = false;
    w = 0;
    // State::new (state.rs:52-81)
    t_curr = t_prev = A.t0;
    double p0[N];
    init_condition_eval<0, N>(A.history_id, A.t0, hist.y0, hist.hist_k, hist.status, p0);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      p_curr[n] = p_prev[n] = p0[n];
      d_curr[n] = 0.0;
    }
    long long s = 0;  // index of the step being taken = ring record it will write
    if (t_curr < A.t1) {
      rhs(d_curr);  // k[0] of the first step (state.rs:97-98)
      const double h = A.h;
#pragma unroll 1
      while (t_curr < A.t1 && (A.t1 - t_curr) >= h && hist.status == 0u) {
        step<true>(h);
        commit<true>(s, h);
        refill(s, tau, h);
        ++s;
      }
#pragma unroll 1
      while (t_curr < A.t1 && hist.status == 0u) {
        const double h_req = fmin(h, A.t1 - t_curr);
        step<false>(h_req);
        commit<false>(s, h_req);
        refill(s, tau, h);
        ++s;
      }
    }
    bool finite = true;
#pragma unroll
    for (int n = 0; n < N; ++n) finite = finite && (fabs(p_curr[n]) < inf);
    uint32_t status = hist.status;
    if (!finite) status |= DFB_TRAJ_NONFINITE;
    A.t_final[gid] = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + gid] = p_curr[n];
    A.steps[gid] = (unsigned long long)s;
    A.rhs_evals[gid] = s > 0 ? (unsigned long long)s * S + 1 : 0;
    A.status[gid] = status;
  }
};

template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, unsigned long long BIMASK, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
    dde_fixed_kernel(const __grid_constant__ DdeFixedArgs A) {
  const size_t gid = size_t(blockIdx.x) * BLOCK + threadIdx.x;
  if (gid >= A.n_traj) return;
  DdeFixedStepper<Sys, S, I, AMASK, BMASK, BIMASK> st(A, gid);
  st.run();
}

}  // namespace DFB_NS
