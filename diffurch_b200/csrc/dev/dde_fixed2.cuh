// Hot path #2, two members per thread.  Same algorithm, ring layout and record format as
// dev/dde_fixed.cuh (read its header first); what changes is the work per thread.
//
// With a fixed step and ONE delay shared by the whole warp, everything that is not arithmetic on
// a member's own state is the same for every member: the time grid, the window position
// (w, tw0, tw1), the replay of the prune, the loop control, the prefetch bookkeeping, and the
// ~50 tableau / interpolant constants that do not fit the 63 uniform registers and are re-fetched
// from the constant bank every step.  Round 1's dde_fixed_kernel paid all of that per member (ncu: 355
// instructions per member-step, 72 % of the issue slots — issue-bound, not HBM-bound, since
// the records shrank to 40 B).  Here a thread integrates TPT members (trajectory g and
// g + grid threads): the shared part is paid once per TPT members and the TPT independent
// dependency chains interleave, as in ode_fixed.cuh.  (Round 2 cut the one-member kernel's shared part
// instead — lean steady loop, ~280 instructions per member-step — and it now wins at most sizes; the
// launcher chooses per ensemble size, see kernels_dde_fixed.cu.)
//
// A warp whose members do not all share one delay (or whose second slot runs past the end of the
// ensemble) is handed, slot by slot, to the one-member stepper of dde_fixed.cuh inside the same
// kernel: results never depend on which path a warp took.
#pragma once
#include "dde_fixed.cuh"

namespace DFB_NS {

template <class Stepper, bool STEADY>
struct WindowHistT {
  Stepper* st;
  int u;  // slot of the thread (compile-time after unrolling)
  template <int D>
  __device__ __forceinline__ void eval(double tt, double* out) {
    constexpr int N = Stepper::N, I = Stepper::I_;
    if (!STEADY) {
      if (tt <= st->A.t0) {  // state.rs:162-163: the history function (cold: state reloaded from HBM)
        double y0[N];
#pragma unroll
        for (int n = 0; n < N; ++n) y0[n] = st->A.y0[size_t(n) * st->A.n_traj + st->gid[u]];
        const double hk = Stepper::SysT::NP >= 4 ? st->A.params[size_t(3) * st->A.n_traj + st->gid[u]] : 0.0;
        init_condition_eval<D, N>(st->A.history_id, tt, y0, hk, st->status[u], out);
        return;
      }
      if (tt >= st->t_prev) st->status[u] |= DFB_TRAJ_HISTORY_NOT_READY;  // state.rs:172-177 (delay <= h)
      if (tt < st->t_oldest) st->status[u] |= DFB_TRAJ_HISTORY_DELETED;   // state.rs:166-171 (max_delay < delay)
    }
    const bool second = tt >= st->tw1;
    const double* p = (second ? st->p1 : st->p0) + u * Stepper::kWarpRec;
    eval_staged<D, N, I>(p, tt - (second ? st->tw1 : st->tw0), out);
  }
};

template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, unsigned long long BIMASK, int TPT>
struct DdeFixedStepperT {
  using SysT = Sys;
  using Self = DdeFixedStepperT;
  static constexpr int N = Sys::N;
  static constexpr int I_ = I;
  static constexpr int NP = Sys::NP > 0 ? Sys::NP : 1;
  static constexpr int F = N * I;          // doubles per record
  static constexpr int kWarpRec = F * 32;  // doubles of one warp's record block (F x 256 bytes)

  const DdeFixedArgs& A;
  int lane;
  double* sm;  // this warp's staging area [kDdeSmemSlots][TPT][F][32]
  size_t gid[TPT];
  bool valid[TPT];
  unsigned warp_base[TPT], slot_stride;  // ring addressing in doubles (host guarantees < 2^32)
  // shared by the TPT members of the thread (and by the warp)
  double t_curr, t_prev;
  double tw0, tw1;
  const double *p0, *p1;  // slot 0's column of the two staged records; slot u is u * kWarpRec further
  double t_oldest, t_old1;
  int w;
  bool loaded, track_oldest;
  // per member
  double p_curr[TPT][N], p_prev[TPT][N], d_curr[TPT][N];
  double k[TPT][S][N];
  double prm[TPT][NP];
  uint32_t status[TPT];
  uint32_t n_samples = 0u, n_samples_alive[TPT] = {};

  __device__ __forceinline__ void on_step(unsigned calls) {  // see DdeFixedStepper::on_step
    if (A.trace_every > 0 && calls % unsigned(A.trace_every) == 0u && n_samples < uint32_t(A.trace_capacity)) {
#pragma unroll
      for (int u = 0; u < TPT; ++u) {
        if (!valid[u]) continue;
        A.trace_t[size_t(n_samples) * A.n_traj + gid[u]] = t_curr;
#pragma unroll
        for (int n = 0; n < N; ++n) A.trace_y[(size_t(n_samples) * N + n) * A.n_traj + gid[u]] = p_curr[u][n];
      }
      ++n_samples;
#pragma unroll
      for (int u = 0; u < TPT; ++u)
        if (!(status[u] & kDdePanicBits)) n_samples_alive[u] = n_samples;
    }
  }

  __device__ DdeFixedStepperT(const DdeFixedArgs& A_, double* sm_) : A(A_), lane(int(threadIdx.x & 31)), sm(sm_) {}

  template <bool STEADY>
  __device__ __forceinline__ void rhs(int u, double* out) {
    WindowHistT<Self, STEADY> h{this, u};
    Sys::rhs(StateRef<N, WindowHistT<Self, STEADY>>{t_curr, p_curr[u], d_curr[u], t_prev, p_prev[u], &h}, prm[u], out);
  }

  __device__ __forceinline__ static bool amask(int i, int j) { return (AMASK >> (i * 8 + j)) & 1ull; }

  template <bool FULL, bool STEADY>
  __device__ __forceinline__ void step(double t_step) {
#pragma unroll
    for (int u = 0; u < TPT; ++u)
#pragma unroll
      for (int n = 0; n < N; ++n) {
        k[u][0][n] = d_curr[u][n];
        p_prev[u][n] = p_curr[u][n];
      }
    t_prev = t_curr;
#pragma unroll
    for (int i = 1; i < S; ++i) {
      t_curr = t_prev + (FULL ? A.hc[i] : A.c[i] * t_step);
#pragma unroll
      for (int u = 0; u < TPT; ++u)
#pragma unroll
        for (int n = 0; n < N; ++n) {
          if (FULL) {
            double acc = p_prev[u][n];
#pragma unroll
            for (int j = 0; j < i; ++j)
              if (amask(i, j)) acc = fma(k[u][j][n], A.ha[i * DFB_MAX_STAGES + j], acc);
            p_curr[u][n] = acc;
          } else {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < i; ++j)
              if (amask(i, j)) acc = fma(k[u][j][n], A.a[i * DFB_MAX_STAGES + j], acc);
            p_curr[u][n] = fma(acc, t_step, p_prev[u][n]);
          }
        }
#pragma unroll
      for (int u = 0; u < TPT; ++u) rhs<STEADY>(u, k[u][i]);
    }
#pragma unroll
    for (int u = 0; u < TPT; ++u)
#pragma unroll
      for (int n = 0; n < N; ++n) {
        if (FULL) {
          double acc = p_prev[u][n];
#pragma unroll
          for (int j = 0; j < S; ++j)
            if (BMASK & (1u << j)) acc = fma(k[u][j][n], A.hb[j], acc);
          p_curr[u][n] = acc;
        } else {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < S; ++j)
            if (BMASK & (1u << j)) acc = fma(k[u][j][n], A.b[j], acc);
          p_curr[u][n] = fma(acc, t_step, p_prev[u][n]);
        }
      }
    t_curr = t_prev + t_step;
#pragma unroll
    for (int u = 0; u < TPT; ++u) rhs<STEADY>(u, d_curr[u]);
  }

  // ---- ring I/O: [slot][warp][field][lane]; staging [slot & 3][u][field][lane] ------------
  __device__ __forceinline__ double* warp_block(int rec, int u) const {
    const unsigned slot = unsigned(rec) & unsigned(A.ring_capacity - 1);
    return A.ring + (slot * slot_stride + warp_base[u]);
  }
  __device__ __forceinline__ double* smem_block(int rec) const {
    return sm + (unsigned(rec) & unsigned(kDdeSmemSlots - 1)) * unsigned(TPT * kWarpRec);
  }
  __device__ __forceinline__ void issue_fetch(int rec) {
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      const double* g = warp_block(rec, u);
      double* d = smem_block(rec) + u * kWarpRec;
#pragma unroll
      for (int q = 0; q < F / 2; ++q) cp_async16(d + 2 * (lane + 32 * q), g + 2 * (lane + 32 * q));
      if ((F & 1) && lane < 16) cp_async16(d + 2 * (lane + 32 * (F / 2)), g + 2 * (lane + 32 * (F / 2)));
    }
  }

  // issue_fetch with the lane's shared-memory address and ring columns computed once (lean steady loop)
  __device__ __forceinline__ void issue_fetch_lean(int rec, unsigned sm_lane, const double* const* ring_lane) {
    const unsigned slot_off = (unsigned(rec) & unsigned(A.ring_capacity - 1)) * slot_stride;
    const unsigned d0 = sm_lane + (unsigned(rec) & unsigned(kDdeSmemSlots - 1)) * unsigned(TPT * kWarpRec * sizeof(double));
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      const double* g = ring_lane[u] + slot_off;
      const unsigned d = d0 + unsigned(u * kWarpRec * sizeof(double));
#pragma unroll
      for (int q = 0; q < F / 2; ++q) cp_async16_saddr(d + 512u * q, g + 64 * q);
      if ((F & 1) && lane < 16) cp_async16_saddr(d + 512u * (F / 2), g + 64 * (F / 2));
    }
  }

  // commit_step (state.rs:122-139): pre-combined interpolant of the step just taken -> ring
  template <bool FULL>
  __device__ __forceinline__ void commit(int s, double t_step) {
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      double inv = 1.0;  // 1 / h^(j-1) for partial steps
      double* p = warp_block(s, u) + lane;
#pragma unroll
      for (int n = 0; n < N; ++n) __stcs(p + 32 * n, p_prev[u][n]);
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
            acc = first ? k[u][i][n] * cf : fma(k[u][i][n], cf, acc);
            first = false;
          }
          __stcs(p + 32 * (N * j + n), FULL ? acc : acc * inv);
        }
        if (!FULL) inv = inv / t_step;
      }
    }
  }

  // Same bookkeeping as DdeFixedStepper::refill (uniform path), once for the TPT members.
  __device__ __forceinline__ void refill(int s, double tau, double h) {
    if (track_oldest) {
      const double t_tail = t_prev - A.max_delay;
      while (t_old1 < t_tail && t_old1 < t_prev) {
        t_oldest = t_old1;
        t_old1 = t_old1 + h;
      }
    }
    int issued_before;
    if (!loaded) {
      if (!(t_curr + h - tau > A.t0)) return;
      loaded = true;
      w = 0;
      tw0 = A.t0;
      tw1 = A.t0 + h;
      issued_before = -1;
    } else {
      issued_before = (w + 2 < s - 1) ? w + 2 : s - 1;
    }
    const double next_len = fmin(h, A.t1 - t_curr);
    const double tt_min = (t_curr + (next_len == h ? A.hc_min : A.c_min * next_len)) - tau;
    while (tw1 <= tt_min) {
      ++w;
      tw0 = tw1;
      tw1 = tw1 + h;
    }
    const int want = (w + 2 < s) ? w + 2 : s;
    if (want >= s) __syncwarp();
#pragma unroll 1
    for (int r = issued_before + 1; r <= want; ++r) issue_fetch(r);
    cp_async_commit();
    const int need = (w + 1 < s) ? w + 1 : s;
    if (need <= issued_before) cp_async_wait<1>();
    else cp_async_wait<0>();
    __syncwarp();
    p0 = smem_block(w) + lane;
    p1 = smem_block(w + 1) + lane;
  }

  // refill for the steady regime (see DdeFixedStepper::refill_steady)
  __device__ __forceinline__ void refill_steady(int s, double tau, double h) {
    if (!(!track_oldest && loaded && w + 4 < s)) {
      refill(s, tau, h);
      return;
    }
    const double next_len = fmin(h, A.t1 - t_curr);
    const double tt_min = (t_curr + (next_len == h ? A.hc_min : A.c_min * next_len)) - tau;
    int r = w + 3;
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
    else cp_async_wait<1>();
    __syncwarp();
    p0 = smem_block(w) + lane;
    p1 = smem_block(w + 1) + lane;
  }

  // gid_[u] / valid_[u] / g_[u] (global thread slot index) set by the kernel; tau is the warp's delay
  __device__ void run(double tau) {
    const size_t n_traj = A.n_traj;
    track_oldest = !(A.max_delay >= tau);
    const size_t n_warps = (n_traj + 31) / 32;
    slot_stride = unsigned(n_warps * size_t(kWarpRec));
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    t_oldest = A.t0;
    t_old1 = A.t0 + A.h;
    tw0 = tw1 = inf;
    p0 = p1 = sm + lane;
    loaded = false;
    w = 0;
    t_curr = t_prev = A.t0;  // State::new (state.rs:52-81)
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      status[u] = 0u;
#pragma unroll
      for (int n = 0; n < Sys::NP; ++n) prm[u][n] = A.params[size_t(n) * n_traj + gid[u]];
      double y0[N], p0v[N];
#pragma unroll
      for (int n = 0; n < N; ++n) y0[n] = A.y0[size_t(n) * n_traj + gid[u]];
      const double hk = Sys::NP >= 4 ? prm[u][Sys::NP >= 4 ? 3 : 0] : 0.0;
      init_condition_eval<0, N>(A.history_id, A.t0, y0, hk, status[u], p0v);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        p_curr[u][n] = p_prev[u][n] = p0v[n];
        d_curr[u][n] = 0.0;
      }
    }
    int s = 0;
    on_step(0u);  // events_on_step at t_init (solver.rs:290)
    if (t_curr < A.t1) {
#pragma unroll
      for (int u = 0; u < TPT; ++u) rhs<false>(u, d_curr[u]);  // k[0] of the first step (state.rs:97-98)
      const double h = A.h;
      const bool counted = A.n_full_steps >= 0;
      const int n_full = int(A.n_full_steps);
      // Phase A: lookups may still reach the history function / the delay does not exceed the step
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
      // Phase B, lean form (see DdeFixedStepper::run): counted full steps, no prune replay, no trace
      if (counted && !track_oldest && loaded && A.trace_every <= 0) {
        const unsigned sm_lane = (unsigned)__cvta_generic_to_shared(sm) + 16u * unsigned(lane);
        const double* ring_lane[TPT];
#pragma unroll
        for (int u = 0; u < TPT; ++u) ring_lane[u] = A.ring + warp_base[u] + 2 * lane;
#pragma unroll 1
        while (s + 1 < n_full && w + 4 < s) {
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
            for (int r = w_was + 3; r < w + 2; ++r) issue_fetch_lean(r, sm_lane, ring_lane);
            issue_fetch_lean(w + 2, sm_lane, ring_lane);
            cp_async_commit();
            cp_async_wait<0>();  // the window jumped: its records may have been issued only now
          } else if (w != w_was) {
            issue_fetch_lean(w + 2, sm_lane, ring_lane);
          }
          cp_async_commit();
          cp_async_wait<1>();
          __syncwarp();
          p0 = smem_block(w) + lane;
          p1 = smem_block(w + 1) + lane;
          ++s;
        }
      }
      // Phase B (steady)
#pragma unroll 1
      while (counted ? s < n_full : (t_curr < A.t1 && (A.t1 - t_curr) >= h)) {
        if (track_oldest && (t_curr + A.hc_min) - tau < t_oldest) {
#pragma unroll
          for (int u = 0; u < TPT; ++u) status[u] |= DFB_TRAJ_HISTORY_DELETED;
        }
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
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      if (!valid[u]) continue;
      bool finite = true;
#pragma unroll
      for (int n = 0; n < N; ++n) finite = finite && (fabs(p_curr[u][n]) < inf);
      if (!finite && status[u] == 0u) status[u] |= DFB_TRAJ_NONFINITE;
      A.t_final[gid[u]] = t_curr;
#pragma unroll
      for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + gid[u]] = p_curr[u][n];
      A.steps[gid[u]] = (unsigned long long)s;
      A.rhs_evals[gid[u]] = s > 0 ? (unsigned long long)s * S + 1 : 0;
      A.status[gid[u]] = status[u];
      if (A.trace_every > 0 && A.n_samples) A.n_samples[gid[u]] = n_samples_alive[u];
    }
  }
};

template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, unsigned long long BIMASK,
          int BLOCK, int MIN_BLOCKS>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
    dde_fixed2_kernel(const __grid_constant__ DdeFixedArgs A) {
  constexpr int TPT = 2;
  using StepperT = DdeFixedStepperT<Sys, S, I, AMASK, BMASK, BIMASK, TPT>;
  using Stepper1 = DdeFixedStepper<Sys, S, I, AMASK, BMASK, BIMASK>;
  extern __shared__ __align__(16) double smem[];  // (blockDim.x / 32) * kDdeSmemSlots * TPT * kWarpRec doubles
  double* sm_warp = smem + (threadIdx.x / 32) * kDdeSmemSlots * TPT * StepperT::kWarpRec;
  const size_t g0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  if ((g0 & ~size_t(31)) >= A.n_traj) return;  // whole warp past the end in slot 0 (and then in slot 1 too)
  size_t g[TPT], gid[TPT];
  bool valid[TPT];
  double tau[TPT];
#pragma unroll
  for (int u = 0; u < TPT; ++u) {
    g[u] = g0 + size_t(u) * stride;
    valid[u] = g[u] < A.n_traj;
    gid[u] = valid[u] ? g[u] : A.n_traj - 1;
    double prm[Sys::NP > 0 ? Sys::NP : 1];
#pragma unroll
    for (int n = 0; n < Sys::NP; ++n) prm[n] = A.params[size_t(n) * A.n_traj + gid[u]];
    tau[u] = Sys::const_delay(prm);
  }
  const bool slot1_in_range = (g[1] & ~size_t(31)) < A.n_traj;  // warp-uniform
  const long long tau0 = __shfl_sync(0xffffffffu, __double_as_longlong(tau[0]), 0);
  const bool same = __double_as_longlong(tau[0]) == tau0 && __double_as_longlong(tau[1]) == tau0;
  const bool uniform = slot1_in_range && __all_sync(0xffffffffu, same);
  if (uniform) {
    StepperT st(A, sm_warp);
#pragma unroll
    for (int u = 0; u < TPT; ++u) {
      st.gid[u] = gid[u];
      st.valid[u] = valid[u];
      st.warp_base[u] = unsigned((g[u] / 32) * size_t(StepperT::kWarpRec));
    }
    st.run(__longlong_as_double(tau0));
  } else {
    // members with different delays in one warp (or a ragged tail): one member at a time through
    // the stepper of dde_fixed.cuh, which also has the per-lane window path
#pragma unroll 1
    for (int u = 0; u < TPT; ++u) {
      if (u == 1 && !slot1_in_range) break;
      Stepper1 st(A, g[u], gid[u], valid[u], sm_warp);
      st.run();
      __syncwarp();
    }
  }
}

}  // namespace DFB_NS
