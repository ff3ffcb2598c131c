// Launch geometry of the one-member-per-lane hot kernels, chosen from the ENSEMBLE SIZE.
//
// The kernels integrate a fixed set of members per thread for the whole launch, so for a pipe-bound
// kernel the launch lasts as long as the busiest SM sub-partition (SMSP: own warp scheduler and own
// FP64 lanes) has work.  A shard of the BASELINE ensemble on one of 8 GPUs is small enough for that
// quantisation to matter: 131 072 Lorenz members at 2 members/thread are 2 048 warps = 3.46 per SMSP,
// so the busiest SMSP carries 4 (86 %); at 1 member/thread they are 6.92 per SMSP -> 7 (98.8 %).
// `choose_geometry` scores every (members per thread, block size) candidate by the members the
// busiest SMSP would integrate and takes the cheapest; large ensembles keep 2 members per thread
// (fewer non-FP64 instructions per member), small ones spread out.
#pragma once
#include <cstddef>
#include <cstdlib>

namespace dfb_host {

struct LaunchGeometry {
  int block = 128;
  int tpt = 1;
  size_t grid = 1;
  double cost = 0.0;         // members (x per-member weight) on the busiest SM sub-partition
  double lower_bound = 0.0;  // n / (SMs x 4 SMSPs x 32 lanes)
};

struct GeometryCandidate {
  int block;
  int tpt;
  int resident_blocks;  // cudaOccupancyMaxActiveBlocksPerMultiprocessor for this (kernel, block)
  double weight;        // relative cost of one member-step in this variant (2 members/thread = 1.0)
};

inline double geometry_cost(size_t n, int n_sm, const GeometryCandidate& c) {
  const size_t per_block = size_t(c.block) * size_t(c.tpt);
  const size_t ctas = (n + per_block - 1) / per_block;
  const size_t warps_per_cta = size_t(c.block + 31) / 32;
  // Blocks are handed out evenly, so the busiest SM integrates ceil(ctas / SMs) of them and its busiest
  // sub-partition a quarter of their warps, rounded up — whether they are resident together (one wave)
  // or one after another (several waves: a pipe-bound block runs faster while fewer share the pipe, so
  // the launch still lasts as long as the SM has work).  Measured on B200 (profiles/r2_shards_*.json):
  // 2^20 / 2^19 / 2^18 members favour 2 per thread by 1.2 %, 2^17 favours 1 per thread by 11 %.
  const size_t c_sm = (ctas + size_t(n_sm) - 1) / size_t(n_sm);
  const size_t busiest_warps = (c_sm * warps_per_cta + 3) / 4;
  (void)c.resident_blocks;
  return double(busiest_warps) * double(c.tpt) * c.weight;
}

// first candidate wins ties: list the preferred (largest) geometry first
inline LaunchGeometry choose_geometry(size_t n, int n_sm, const GeometryCandidate* cand, int n_cand) {
  LaunchGeometry best;
  if (n_sm <= 0) n_sm = 148;
  best.lower_bound = double(n) / (double(n_sm) * 128.0);
  bool have = false;
  for (int i = 0; i < n_cand; ++i) {
    const double cost = geometry_cost(n, n_sm, cand[i]);
    if (!have || cost < best.cost * (1.0 - 1e-9)) {
      have = true;
      best.block = cand[i].block;
      best.tpt = cand[i].tpt;
      best.cost = cost;
      const size_t per_block = size_t(cand[i].block) * size_t(cand[i].tpt);
      best.grid = (n + per_block - 1) / per_block;
    }
  }
  return best;
}

inline int env_int(const char* name, int fallback) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : fallback;
}

}  // namespace dfb_host
