// Instantiations + launcher of the constant-delay DDE hot-path kernel (FAST
// arithmetic only; STRICT goes through the general integrator, which keeps the
// reference's record format and operation order).
#include "dev/dde_fixed2.cuh"

#include <cmath>
#include <cstring>

namespace DFB_NS {
namespace {
constexpr int kBlock = 128;
#ifndef DFB_DDE_MINB2
#define DFB_DDE_MINB2 4
#endif
constexpr int kMinBlocks2 = DFB_DDE_MINB2;  // two members per thread: <= 128 registers at 4 blocks/SM
constexpr int kMinBlocks = 7;  // 7 x 128 threads: <= 72 registers; 2^18 members = 1.98 full waves on 148 SMs

constexpr unsigned long long bit8(int i, int j) { return 1ull << (i * 8 + j); }
constexpr unsigned long long full_lower(int S) {
  unsigned long long m = 0;
  for (int i = 1; i < S; ++i)
    for (int j = 0; j < i; ++j) m |= bit8(i, j);
  return m;
}
constexpr unsigned long long full_bi(int S, int I) {
  unsigned long long m = 0;
  for (int i = 0; i < S; ++i)
    for (int j = 1; j < I; ++j) m |= bit8(i, j);
  return m;
}
// rktp64 (rk.rs:313-429): row 0 -> theta^1..4, rows 2,3,4,6 -> theta^2..4, row 5 -> theta^4, row 1 zero
constexpr unsigned long long kBiRktp64 =
    bit8(0, 1) | bit8(0, 2) | bit8(0, 3) | bit8(0, 4) | bit8(2, 2) | bit8(2, 3) | bit8(2, 4) |
    bit8(3, 2) | bit8(3, 3) | bit8(3, 4) | bit8(4, 2) | bit8(4, 3) | bit8(4, 4) | bit8(5, 4) |
    bit8(6, 2) | bit8(6, 3) | bit8(6, 4);

template <class Sys, int S, int I, unsigned long long AMASK, unsigned BMASK, unsigned long long BIMASK>
bool try_launch(int S_rt, int I_rt, unsigned long long am, unsigned bm, unsigned long long bim,
                const DdeFixedArgs& A, cudaStream_t stream, int* rc) {
  if (S_rt != S || I_rt != I) return false;
  if ((am & ~AMASK) || (bm & ~BMASK) || (bim & ~BIMASK)) return false;
  // Members per thread and block size from the ensemble size.  The kernel is HBM-bound once the device is
  // full and latency-bound below that (a 2^18-member ensemble sharded over 8 GPUs is 1 024 warps for 592 SM
  // sub-partitions).  Two builds: one member per thread (dev/dde_fixed.cuh, PRE: delayed values ahead of the
  // stage chain, lean steady loop, 8 staging slots) and two per thread (dev/dde_fixed2.cuh); the wave/tail
  // model below picks.  Small ensembles are spread over every SM in 32- or 64-thread blocks.
  // DFB_DDE_TPT=1|2 / DFB_DDE_BLOCK / DFB_DDE_RELAXED override (tuning, tests).
  using Stepper = DdeFixedStepper<Sys, S, I, AMASK, BMASK, BIMASK>;
  auto kernel2 = dde_fixed2_kernel<Sys, S, I, AMASK, BMASK, BIMASK, kBlock, kMinBlocks2>;
  auto kernel1 = dde_fixed_kernel<Sys, S, I, AMASK, BMASK, BIMASK, kBlock, kMinBlocks>;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(kernel2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(kernel1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         int((kBlock / 32) * kDdeSmemSlots * 2 * Stepper::kWarpRec * sizeof(double)));
    cudaFuncSetAttribute(kernel1, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         int((kBlock / 32) * kDdeSmemSlots * Stepper::kWarpRec * sizeof(double)));
  }
  const size_t warps = (A.n_traj + 31) / 32;
  // below ~16 warps per SM one warp per block spreads the ensemble over every SM sub-partition
  int block = warps >= size_t(n_sm) * 16 ? kBlock : (warps >= size_t(n_sm) * 8 ? 64 : 32);
  // One or two members per thread: both builds keep 16 warps per SM resident, so a launch is
  // floor(waves) full waves at the HBM-bound rate plus a tail that cannot finish faster than a lone warp
  // runs its own dependent instruction stream (measured: ~0.48 of a full wave's time with one member per
  // thread, ~0.33 with two).  The cheaper prediction wins; two members per thread carry a 3 % handicap
  // (measured at equal predictions).  2^18 members: one per thread 17.9 ms vs 18.3; 3/8 of it: two per
  // thread 6.6 vs 7.5 ms; 1/8: one per thread 2.55 vs 3.4 ms (profiles/r2_dde_geometry_model.txt).
  auto predicted = [&](int members_per_thread) {
    const double slots = double(n_sm) * 16.0;  // resident warps on the device
    const double waves = double((A.n_traj + 32 * members_per_thread - 1) / (32 * members_per_thread)) / slots;
    const double t_full = slots * 32.0 * members_per_thread;
    const double t_min = (members_per_thread == 1 ? 0.48 : 0.33) * t_full;
    const double whole = double(size_t(waves));
    const double frac = waves - whole;
    double t = whole * t_full;
    if (frac > 0.0) t += (frac * t_full > t_min) ? frac * t_full : t_min;
    return members_per_thread == 1 ? t : 1.03 * t;
  };
  int tpt = predicted(2) < predicted(1) ? 2 : 1;
  if (const char* e = getenv("DFB_DDE_TPT")) tpt = atoi(e) == 2 ? 2 : 1;
  if (const char* e = getenv("DFB_DDE_BLOCK")) block = atoi(e) == 32 ? 32 : (atoi(e) == 64 ? 64 : kBlock);
  if (tpt == 2) {
    const size_t grid = (A.n_traj + size_t(block) * 2 - 1) / (size_t(block) * 2);
    const size_t smem = size_t(block / 32) * kDdeSmemSlots * 2 * Stepper::kWarpRec * sizeof(double);
    kernel2<<<unsigned(grid), block, smem, stream>>>(A);
  } else {
    const size_t grid = (A.n_traj + block - 1) / block;
    size_t smem = size_t(block / 32) * kDdeSmemSlots * Stepper::kWarpRec * sizeof(double);
    // Few warps per SM (a shard of the ensemble): nothing hides the latency of a warp's own instruction
    // stream, so the build WITHOUT the 72-register cap runs — the scheduler may then hoist the step's
    // state-independent history evaluations above the dependent stage chain.  DFB_DDE_RELAXED=0|1 overrides.
    // Measured faster at every size the one-member kernel is chosen for (2^16: 5.18 vs 5.58 ms).
    bool relaxed = true;
    if (const char* e = getenv("DFB_DDE_RELAXED")) relaxed = atoi(e) != 0;
    if (relaxed) {
      auto kernel1r = dde_fixed_kernel<Sys, S, I, AMASK, BMASK, BIMASK, kBlock, 2, true>;
      using StepperPre = DdeFixedStepper<Sys, S, I, AMASK, BMASK, BIMASK, true>;
      smem = size_t(block / 32) * StepperPre::kSlots * Stepper::kWarpRec * sizeof(double);
      static bool once = false;
      if (!once) {
        cudaFuncSetAttribute(kernel1r, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(kernel1r, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int((kBlock / 32) * StepperPre::kSlots * Stepper::kWarpRec * sizeof(double)));
        once = true;
      }
      kernel1r<<<unsigned(grid), block, smem, stream>>>(A);
    } else {
      kernel1<<<unsigned(grid), block, smem, stream>>>(A);
    }
  }
  *rc = int(cudaGetLastError());
  return true;
}

template <class Sys>
bool launch_sys(int S, int I, unsigned long long am, unsigned bm, unsigned long long bim,
                const DdeFixedArgs& A, cudaStream_t st, int* rc) {
  return try_launch<Sys, 7, 5, full_lower(7), 0x7Du, kBiRktp64>(S, I, am, bm, bim, A, st, rc) ||
         try_launch<Sys, 7, 5, full_lower(7), 0x7Fu, full_bi(7, 5)>(S, I, am, bm, bim, A, st, rc) ||
         try_launch<Sys, 5, 4, full_lower(5), 0x1Fu, full_bi(5, 4)>(S, I, am, bm, bim, A, st, rc) ||
         try_launch<Sys, 4, 2, full_lower(4), 0xFu, full_bi(4, 2)>(S, I, am, bm, bim, A, st, rc);
}
}  // namespace

// Doubles of ring the kernel needs per trajectory for (N, I, capacity): N I per record.
size_t dde_fixed_ring_doubles(int n_state, int I, int capacity) {
  return size_t(capacity) * size_t(n_state * I);
}

// returns 0 = launched, >0 = CUDA error code, -1 = no compiled kernel
int run_dde_fixed_shim(int system_id, const dfb_tableau& rk, double t0, double t1, double h,
                       double max_delay, int history_id, int ring_capacity, size_t n,
                       const double* y0, const double* params, double* ring, double* t_final,
                       double* y_final, unsigned long long* steps, unsigned long long* rhs_evals,
                       uint32_t* status, int trace_every, int trace_capacity, double* trace_t,
                       double* trace_y, uint32_t* n_samples, cudaStream_t stream) {
  DdeFixedArgs A;
  std::memset(&A, 0, sizeof(A));
  unsigned long long am = 0, bim = 0;
  unsigned bm = 0;
  for (int i = 0; i < DFB_MAX_STAGES; ++i) {
    for (int j = 0; j < DFB_MAX_STAGES; ++j) {
      const double a = rk.a[i * DFB_MAX_STAGES + j];
      A.a[i * DFB_MAX_STAGES + j] = a;
      A.ha[i * DFB_MAX_STAGES + j] = h * a;
      if (i < rk.S && j < i && a != 0.0) am |= bit8(i, j);
    }
    A.b[i] = rk.b[i];
    A.c[i] = rk.c[i];
    A.hb[i] = h * rk.b[i];
    A.hc[i] = rk.c[i] * h;
    if (i < rk.S && rk.b[i] != 0.0) bm |= 1u << i;
    double hp = 1.0;  // h^(j-1)
    for (int j = 0; j < DFB_MAX_INTERP; ++j) {
      const double v = rk.bi[i * DFB_MAX_INTERP + j];
      A.bi[i * DFB_MAX_INTERP + j] = v;
      if (j >= 1) {
        A.bih[i * DFB_MAX_INTERP + j] = v / hp;
        hp *= h;
        if (i < rk.S && j < rk.I && v != 0.0) bim |= bit8(i, j);
      } else if (i < rk.S && v != 0.0) {
        return -1;  // b_i(0) != 0: not a continuous extension this kernel can pre-combine
      }
    }
  }
  A.t0 = t0;
  A.t1 = t1;
  A.h = h;
  A.max_delay = max_delay;
  A.hc_min = A.hc[rk.S > 1 ? 1 : 0];
  A.c_min = A.c[rk.S > 1 ? 1 : 0];
  for (int i = 1; i < rk.S; ++i)
    if (A.hc[i] < A.hc_min) {
      A.hc_min = A.hc[i];
      A.c_min = A.c[i];
    }
  if (!(A.c_min > 0.0)) return -1;  // a stage at the step start would look up t_prev - tau: general kernel
  A.history_id = history_id;
  A.ring_capacity = ring_capacity;
  A.n_full_steps = -1;
  if (h > 0.0 && (t1 - t0) / h < 5.0e7) {  // replay of solver.rs:292-293 on the shared time grid
    long long cnt = 0;
    volatile double t = t0;                 // volatile: one IEEE double addition per step, as on the device
    while (t < t1 && (t1 - t) >= h) {
      t = t + h;
      ++cnt;
    }
    if (cnt < 2000000000LL) A.n_full_steps = cnt;
  }
  A.n_traj = n;
  A.y0 = y0;
  A.params = params;
  A.t_final = t_final;
  A.y_final = y_final;
  A.steps = steps;
  A.rhs_evals = rhs_evals;
  A.status = status;
  A.ring = ring;
  A.trace_every = (trace_t && trace_capacity > 0) ? trace_every : 0;
  A.trace_capacity = trace_capacity;
  A.trace_t = trace_t;
  A.trace_y = trace_y;
  A.n_samples = n_samples;
  int rc = 0;
  bool ok = false;
#ifdef DFB_PLUGIN_SYSTEM
  (void)system_id;
  if constexpr (DFB_PLUGIN_SYSTEM::uses_history && DFB_PLUGIN_SYSTEM::has_const_delay)
    ok = launch_sys<DFB_PLUGIN_SYSTEM>(rk.S, rk.I, am, bm, bim, A, stream, &rc);
  return ok ? rc : -1;
#else
  switch (system_id) {
    case DFB_SYS_DDE_LINEAR: ok = launch_sys<SysDdeLinear>(rk.S, rk.I, am, bm, bim, A, stream, &rc); break;
    case DFB_SYS_NDDE_LINEAR: ok = launch_sys<SysNddeLinear>(rk.S, rk.I, am, bm, bim, A, stream, &rc); break;
    default: break;
  }
  return ok ? rc : -1;
#endif
}

}  // namespace DFB_NS
