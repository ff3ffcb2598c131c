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
  template <bool FULL>
  __device__ __forceinline__ void step(double t_step) {
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
      }
#pragma unroll 1
      while (t_curr < A.t1) {
        step<false>(fmin(h, A.t1 - t_curr));
        ++n_steps;
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
    }
  }
};

template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, bool PRESCALED, int BLOCK, int TPT>
__global__ void __launch_bounds__(BLOCK)
    ode_fixed_kernel(const __grid_constant__ OdeFixedArgs A) {
  // slot u of thread g handles trajectory g + u * (threads in the grid): coalesced for every u
  const size_t g = size_t(blockIdx.x) * BLOCK + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * BLOCK;
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
  st.run(gid, valid);
}

}  // namespace DFB_NS
