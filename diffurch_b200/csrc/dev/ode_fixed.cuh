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

template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, bool PRESCALED>
struct OdeFixedStepper {
  static constexpr int N = Sys::N;
  static constexpr int NP = Sys::NP > 0 ? Sys::NP : 1;
  using Ref = StateRef<N, NoHist>;

  const OdeFixedArgs& A;
  double t_curr, t_prev;
  double p_curr[N], p_prev[N], d_curr[N];
  double k[S][N];
  double prm[NP];
  NoHist nohist;

  __device__ explicit OdeFixedStepper(const OdeFixedArgs& A_) : A(A_) {}

  __device__ __forceinline__ void rhs(double* out) {
    Sys::rhs(Ref{t_curr, p_curr, d_curr, t_prev, p_prev, &nohist}, prm, out);
  }

  // One make_step of length t_step; FULL = (t_step == A.h), the only case the
  // pre-scaled rows are valid for.
  template <bool FULL>
  __device__ __forceinline__ void step(double t_step) {
    // k[0] = d_curr (state.rs:95-96); the caller guarantees d_curr is current
#pragma unroll
    for (int n = 0; n < N; ++n) k[0][n] = d_curr[n];
    t_prev = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) p_prev[n] = p_curr[n];
#pragma unroll
    for (int i = 1; i < S; ++i) {
      t_curr = t_prev + (FULL ? A.hc[i] : A.c[i] * t_step);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        if (PRESCALED && FULL) {
          double acc = p_prev[n];
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (AMASK & amask_bit(i, j)) acc = fma(k[j][n], A.ha[i * DFB_MAX_STAGES + j], acc);
          p_curr[n] = acc;
        } else {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (AMASK & amask_bit(i, j)) acc = acc + k[j][n] * A.a[i * DFB_MAX_STAGES + j];
          p_curr[n] = p_prev[n] + acc * t_step;
        }
      }
      rhs(k[i]);
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
      if (PRESCALED && FULL) {
        double acc = p_prev[n];
#pragma unroll
        for (int j = 0; j < S; ++j)
          if (BMASK & (1u << j)) acc = fma(k[j][n], A.hb[j], acc);
        p_curr[n] = acc;
      } else {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < S; ++j)
          if (BMASK & (1u << j)) acc = acc + k[j][n] * A.b[j];
        p_curr[n] = p_prev[n] + acc * t_step;
      }
    }
    t_curr = t_prev + t_step;
    rhs(d_curr);
  }

  __device__ void run(size_t gid) {
    const size_t n_traj = A.n_traj;
#pragma unroll
    for (int n = 0; n < Sys::NP; ++n) prm[n] = A.params[size_t(n) * n_traj + gid];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      p_curr[n] = p_prev[n] = A.y0[size_t(n) * n_traj + gid];
      d_curr[n] = 0.0;
    }
    t_curr = t_prev = A.t0;
    unsigned long long n_steps = 0;
    if (t_curr < A.t1) {
      rhs(d_curr);  // first make_step evaluates k[0] itself (state.rs:97-98)
      const double h = A.h;
      while (t_curr < A.t1) {
        const double h_req = fmin(h, A.t1 - t_curr);
        if (h_req == h) step<true>(h);
        else step<false>(h_req);
        ++n_steps;
      }
    }
    bool finite = true;
#pragma unroll
    for (int n = 0; n < N; ++n) finite = finite && (fabs(p_curr[n]) < __longlong_as_double(0x7ff0000000000000LL));
    A.t_final[gid] = t_curr;
#pragma unroll
    for (int n = 0; n < N; ++n) A.y_final[size_t(n) * n_traj + gid] = p_curr[n];
    A.steps[gid] = n_steps;
    A.rhs_evals[gid] = n_steps > 0 ? n_steps * S + 1 : 0;
    A.status[gid] = finite ? 0u : uint32_t(DFB_TRAJ_NONFINITE);
  }
};

template <class Sys, int S, unsigned long long AMASK, unsigned BMASK, bool PRESCALED, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
    ode_fixed_kernel(const __grid_constant__ OdeFixedArgs A) {
  const size_t gid = size_t(blockIdx.x) * BLOCK + threadIdx.x;
  if (gid >= A.n_traj) return;
  OdeFixedStepper<Sys, S, AMASK, BMASK, PRESCALED> st(A);
  st.run(gid);
}

}  // namespace DFB_NS
