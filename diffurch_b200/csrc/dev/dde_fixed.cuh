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
// into a per-trajectory ring in HBM laid out [slot][warp][field][lane]: every
// access of a warp is one fully coalesced 256-byte row, and one warp's record is a
// contiguous 8(1 + N I) x 32-byte block.  A lookup is then one
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
  double hc_min;          // min over stages i >= 1 of c_i * h (earliest delayed stage time of a step)
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
  double* ring;  // [cap][ceil(n/32)][1 + N*I][32]
};

template <int N, int I>
struct CoeffRec {
  double t;
  double y[N];
  double d[I > 1 ? I - 1 : 1][N];
};

template <int N, int I>
struct WindowData {
  CoeffRec<N, I> r0, r1;
  double t_init, t_commit;  // t_commit = latest committed time (lookups at or after it are "not yet computed")
  double t_oldest;          // t_deque[0] of the reference after its prune (lookups before it are "deleted")
  double y0[N];
  double hist_k;
  int history_id;
  uint32_t status;
};

// Horner evaluation of one record at offset s = t - t_r.
template <int D, int N, int I>
__device__ __forceinline__ void eval_rec(const CoeffRec<N, I>& r, double s, double* out) {
#pragma unroll
  for (int n = 0; n < N; ++n) {
    if (D == 0) {
      double acc = r.d[I - 2][n];
#pragma unroll
      for (int j = I - 3; j >= 0; --j) acc = fma(acc, s, r.d[j][n]);
      out[n] = fma(acc, s, r.y[n]);
    } else {
      double acc = double(I - 1) * r.d[I - 2][n];
#pragma unroll
      for (int j = I - 3; j >= 0; --j) acc = fma(acc, s, double(j + 1) * r.d[j][n]);
      out[n] = acc;
    }
  }
}

// The two-record window as the `history` of a StateRef.  STEADY = every lookup of
// the step is known to fall after t_init and inside the retained window (checked once
// per step by the stepper), so a lookup is a compare, a subtract and a Horner chain.
template <int N, int I, bool STEADY>
struct WindowHist {
  WindowData<N, I>* w;
  template <int D>
  __device__ __forceinline__ void eval(double tt, double* out) {
    if (!STEADY) {
      if (tt <= w->t_init) {  // state.rs:162-163
        init_condition_eval<D, N>(w->history_id, tt, w->y0, w->hist_k, w->status, out);
        return;
      }
      if (tt >= w->t_commit) w->status |= DFB_TRAJ_HISTORY_NOT_READY;  // state.rs:172-177 (delay <= h)
      if (tt < w->t_oldest) w->status |= DFB_TRAJ_HISTORY_DELETED;     // state.rs:166-171 (max_delay < delay)
    }
    // the branch is warp-uniform for a shared time grid and a shared delay
    if (tt >= w->r1.t) eval_rec<D, N, I>(w->r1, tt - w->r1.t, out);
    else eval_rec<D, N, I>(w->r0, tt - w->r0.t, out);
  }
};

// ---- cp.async (LDGSTS) helpers: global -> shared without staging registers; completion is
// tracked by async groups, so an in-flight prefetch never blocks unrelated instructions.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}

constexpr int kDdeSmemSlots = 4;  // records w .. w+2 of the window plus one being overwritten

template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, unsigned long long BIMASK>
struct DdeFixedStepper {
  static constexpr int N = Sys::N;
  static constexpr int NP = Sys::NP > 0 ? Sys::NP : 1;
  static constexpr int F = 1 + N * I;     // doubles per record
  static constexpr int FP = F + (F & 1);  // padded to an even count: 16-byte cp.async chunks
  static constexpr int kWarpRec = FP * 32;  // doubles of one warp's record block
  using Rec = CoeffRec<N, I>;

  const DdeFixedArgs& A;
  size_t gid;
  bool valid;
  int lane;
  double* sm;  // this warp's staging area [kDdeSmemSlots][FP][32]
  double t_curr, t_prev;
  double p_curr[N], p_prev[N], d_curr[N];
  double k[S][N];
  double prm[NP];
  WindowData<N, I> win;
  double t_old1;       // t_deque[1] of the reference (successor of win.t_oldest on the shared time grid)
  double tw1;          // start time of record w+1, regenerated as t_{r+1} = t_r + h (bit-exact grid replay)
  int w;               // interval index held in win.r0
  int fetched;         // highest record whose copy into shared memory has been issued
  bool loaded, uniform;
  unsigned warp_base, slot_stride;  // ring addressing in doubles: [slot][warp][FP][32] (host guarantees < 2^32)

  size_t g_thread;     // global thread index (= trajectory index for valid lanes)

  __device__ DdeFixedStepper(const DdeFixedArgs& A_, size_t g_, size_t gid_, bool valid_, double* sm_)
      : A(A_), gid(gid_), valid(valid_), lane(int(threadIdx.x & 31)), sm(sm_), g_thread(g_) {}

  template <bool STEADY>
  __device__ __forceinline__ void rhs(double* out) {
    WindowHist<N, I, STEADY> h{&win};
    Sys::rhs(StateRef<N, WindowHist<N, I, STEADY>>{t_curr, p_curr, d_curr, t_prev, p_prev, &h}, prm, out);
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
      rhs<STEADY>(k[i]);
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
    rhs<STEADY>(d_curr);
  }

  // ---- ring I/O: [slot][warp][field][lane] -----------------------------------
  // One warp's record is a contiguous FP x 256-byte block: every field access is one
  // fully coalesced row, field offsets are compile-time immediates, and the block can
  // be moved to shared memory with FP/2 16-byte cp.async per lane.
  __device__ __forceinline__ double* warp_block(int rec) const {
    const unsigned slot = unsigned(rec) & unsigned(A.ring_capacity - 1);
    return A.ring + (slot * slot_stride + warp_base);
  }
  __device__ __forceinline__ double* smem_block(int rec) const {
    return sm + (unsigned(rec) & unsigned(kDdeSmemSlots - 1)) * unsigned(kWarpRec);
  }
  template <class P>
  __device__ __forceinline__ static Rec read_rec(const P* p) {  // p = block base + lane
    Rec r;
    r.t = p[0];
#pragma unroll
    for (int n = 0; n < N; ++n) r.y[n] = p[32 * (1 + n)];
#pragma unroll
    for (int j = 0; j < I - 1; ++j)
#pragma unroll
      for (int n = 0; n < N; ++n) r.d[j][n] = p[32 * (1 + N * (j + 1) + n)];
    return r;
  }
  __device__ __forceinline__ static Rec inf_rec() {
    Rec r;
    r.t = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      r.y[n] = 0.0;
#pragma unroll
      for (int j = 0; j < I - 1; ++j) r.d[j][n] = 0.0;
    }
    return r;
  }
  __device__ __forceinline__ void issue_fetch(int rec) {
    const double* g = warp_block(rec);
    double* d = smem_block(rec);
#pragma unroll
    for (int q = 0; q < FP / 2; ++q) cp_async16(d + 2 * (lane + 32 * q), g + 2 * (lane + 32 * q));
  }

  // commit_step (state.rs:122-139): pre-combine the interpolant of the step just
  // taken and stream it to the ring.  The ring overwrites its oldest slot, which is
  // the prune rule for a window of max_delay.
  template <bool FULL>
  __device__ __forceinline__ void commit(int s, double t_step) {
    double inv = 1.0;  // 1 / h^(j-1) for partial steps
    double* p = warp_block(s) + lane;
    __stcs(p, t_prev);
#pragma unroll
    for (int n = 0; n < N; ++n) __stcs(p + 32 * (1 + n), p_prev[n]);
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
        __stcs(p + 32 * (1 + N * j + n), FULL ? acc : acc * inv);
      }
      if (!FULL) inv = inv / t_step;
    }
  }

  // After committing record s: replay the reference's prune, slide the window so that
  // (r0, r1) bracket the delayed times of the NEXT step, and prefetch one record ahead.
  __device__ __forceinline__ void refill(int s, double tau, double h) {
    win.t_commit = t_curr;
    // prune rule of commit_step (state.rs:126-133) replayed on the shared time grid:
    // pop while t_deque[1] < t_prev - max_delay.  t_{r+1} = t_r + h reproduces the grid bit for bit.
    const double t_tail = t_prev - A.max_delay;
    while (t_old1 < t_tail && t_old1 < t_prev) {
      win.t_oldest = t_old1;
      t_old1 = t_old1 + h;
    }
    if (!loaded) {
      if (!(t_curr + h - tau > A.t0)) return;  // every lookup of the next step is in the history function
      loaded = true;
      w = 0;
      tw1 = A.t0 + h;
    }
    const double tt_min = t_curr - tau;  // earliest delayed time of the next step (c_0 = 0)
    while (tw1 <= tt_min) {  // record w+1 starts at or before it: slide
      ++w;
      tw1 = tw1 + h;
    }
    if (uniform) {
      // stage records up to w+2 (those that exist) in shared memory; the newest one
      // issued here is not needed before the step after next
      const int want = (w + 2 < s) ? w + 2 : s;
      const int fetched_before = fetched;
#pragma unroll 1
      while (fetched < want) issue_fetch(++fetched);
      cp_async_commit();
      const int need = (w + 1 < s) ? w + 1 : s;  // newest record the next step reads
      if (need <= fetched_before) cp_async_wait<1>();   // issued in an earlier refill
      else cp_async_wait<0>();                           // start-up / short delays: issued just now
      __syncwarp();
      win.r0 = read_rec(smem_block(w) + lane);
      win.r1 = (w + 1 <= s) ? read_rec(smem_block(w + 1) + lane) : inf_rec();
    } else {
      // lanes of this warp have different delays: per-lane records straight from HBM
      win.r0 = read_rec(warp_block(w) + lane);
      win.r1 = (w + 1 <= s) ? read_rec(warp_block(w + 1) + lane) : inf_rec();
    }
  }

  __device__ void run() {
    const size_t n_traj = A.n_traj;
#pragma unroll
    for (int n = 0; n < Sys::NP; ++n) prm[n] = A.params[size_t(n) * n_traj + gid];
    const double tau = Sys::const_delay(prm);
    // the staged (warp-cooperative) path needs one window position per warp
    uniform = __all_sync(0xffffffffu, __double_as_longlong(tau) ==
                                          __shfl_sync(0xffffffffu, __double_as_longlong(tau), 0));
    const size_t n_warps = (n_traj + 31) / 32;
    slot_stride = unsigned(n_warps * size_t(kWarpRec));
    warp_base = unsigned((g_thread / 32) * size_t(kWarpRec));
    win.t_init = A.t0;
    win.t_commit = A.t0;
    win.t_oldest = A.t0;
    t_old1 = A.t0 + A.h;
    tw1 = A.t0 + A.h;
    win.history_id = A.history_id;
    win.hist_k = Sys::NP >= 4 ? prm[Sys::NP >= 4 ? 3 : 0] : 0.0;
    win.status = 0u;
#pragma unroll
    for (int n = 0; n < N; ++n) win.y0[n] = A.y0[size_t(n) * n_traj + gid];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    win.r0 = inf_rec();
    win.r1 = inf_rec();
    loaded = false;
    w = 0;
    fetched = -1;
    // State::new (state.rs:52-81)
    t_curr = t_prev = A.t0;
    double p0[N];
    init_condition_eval<0, N>(A.history_id, A.t0, win.y0, win.hist_k, win.status, p0);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      p_curr[n] = p_prev[n] = p0[n];
      d_curr[n] = 0.0;
    }
    int s = 0;  // index of the step being taken = ring record it will write
    // A trajectory that hits one of the reference's panics keeps stepping (its warp must
    // stay converged for the cooperative staging); its status word records the panic.
    if (t_curr < A.t1) {
      rhs<false>(d_curr);  // k[0] of the first step (state.rs:97-98)
      const double h = A.h;
      // Phase A: some lookup of the step may still reach the history function, or
      // the window is not in its steady state: every lookup carries the checks.
#pragma unroll 1
      while (t_curr < A.t1 && (A.t1 - t_curr) >= h &&
             !(loaded && (t_curr - tau) > A.t0 && (t_curr - tau) >= win.r0.t &&
               (t_curr + h) - tau < t_curr)) {
        step<true, false>(h);
        commit<true>(s, h);
        refill(s, tau, h);
        ++s;
      }
      // Phase B (steady): full steps whose lookups all lie after t_init and before the
      // last committed time.  What the reference checks on every lookup is checked once
      // per step: the earliest delayed stage time against the oldest retained record
      // (state.rs:166-171).
#pragma unroll 1
      while (t_curr < A.t1 && (A.t1 - t_curr) >= h) {
        if ((t_curr + A.hc_min) - tau < win.t_oldest) win.status |= DFB_TRAJ_HISTORY_DELETED;
        step<true, true>(h);
        commit<true>(s, h);
        refill(s, tau, h);
        ++s;
      }
      // closing partial step(s): checked path
#pragma unroll 1
      while (t_curr < A.t1) {
        const double h_req = fmin(h, A.t1 - t_curr);
        step<false, false>(h_req);
        commit<false>(s, h_req);
        refill(s, tau, h);
        ++s;
      }
    }
    cp_async_wait<0>();
    if (!valid) return;
    bool finite = true;
#pragma unroll
    for (int n = 0; n < N; ++n) finite = finite && (fabs(p_curr[n]) < inf);
    uint32_t status = win.status;
    if (!finite && status == 0u) status |= DFB_TRAJ_NONFINITE;
    A.t_final[gid] = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + gid] = p_curr[n];
    A.steps[gid] = (unsigned long long)s;
    A.rhs_evals[gid] = s > 0 ? (unsigned long long)s * S + 1 : 0;
    A.status[gid] = status;
  }
};

template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, unsigned long long BIMASK,
          int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
    dde_fixed_kernel(const __grid_constant__ DdeFixedArgs A) {
  using Stepper = DdeFixedStepper<Sys, S, I, AMASK, BMASK, BIMASK>;
  __shared__ __align__(16) double smem[(BLOCK / 32) * kDdeSmemSlots * Stepper::kWarpRec];
  const size_t g = size_t(blockIdx.x) * BLOCK + threadIdx.x;
  // every lane of a warp stays in the kernel (warp-cooperative staging); lanes past the
  // end replay the last trajectory and store nothing.  The ring is allocated for whole warps.
  if ((g & ~size_t(31)) >= A.n_traj) return;  // whole warp past the end (warp-uniform exit)
  const bool valid = g < A.n_traj;
  const size_t gid = valid ? g : A.n_traj - 1;
  Stepper st(A, g, gid, valid, smem + (threadIdx.x / 32) * kDdeSmemSlots * Stepper::kWarpRec);
  st.run();
}

}  // namespace DFB_NS
