// K1 as a stand-alone kernel: one pass of the scan core over the N x S matrix, one candidate per
// warp.  Used by the OMP selection step and by the launch-per-iteration engine (BCG_ENGINE=v1);
// GIGA / Frank-Wolfe builds normally run inside the persistent kernel of loop_kernel.cuh.
// HBM-bound: algorithmic traffic 4*N*S bytes per launch, 0.5-1 flop/byte.
#pragma once
#include "scan_core.cuh"

namespace bcg {

template <int CH, int NDIR, int LPR, int R>
__global__ void __launch_bounds__(512, 1) scan_kernel(const ScanArgs a) {
  using Core = ScanCore<CH, NDIR, LPR, R>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * wpb + warp;
  const int64_t GW = (int64_t)gridDim.x * wpb;
  if (*a.skip0 != 0 || *a.skip1 != 0) return;

  const ScanGeom& q = a.g;
  const uint32_t stage_floats = (uint32_t)q.rps * (uint32_t)q.ld;
  float* wbuf = reinterpret_cast<float*>(smem_raw) + (size_t)warp * q.stages * stage_floats;
  uint64_t* bars =
      reinterpret_cast<uint64_t*>(smem_raw + (size_t)wpb * q.stages * stage_floats * sizeof(float)) + warp * q.stages;
  if (lane == 0) {
    for (int s = 0; s < q.stages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint64_t policy = l2_policy(q.evict_first);

  const int g = lane & (LPR - 1);
  const int grp = lane / LPR;
  const int nchunk = q.ld >> 2;
  float4 d0[CH];
  float4 d1[CH];
  Core::load_dirs(a.dir, q.ld, nchunk, g, d0, d1);

  const int64_t n_chunks = (q.n_rows + q.rps - 1) / q.rps;
  const int64_t n_my = (gw < n_chunks) ? (n_chunks - gw + GW - 1) / GW : 0;
  const int64_t row_step = GW * q.rps;
  // slot s of the ring always holds chunk k with k % stages == s, so a refill goes into the slot
  // that was just consumed; all ring bookkeeping is incremental (no 64-bit div/mod in the loop)
  const bool leader = lane == 0;
  auto issue = [&](int slot, int64_t row0) {        // executed by all lanes, issued by the leader only
    const int64_t left = q.n_rows - row0;
    const uint32_t nr = (uint32_t)(left < q.rps ? left : q.rps);
    tma_load_rows(&bars[slot], wbuf + (size_t)slot * stage_floats, q.An + (size_t)row0 * q.ld, nr * (uint32_t)q.ld * 4u,
                  policy, leader);
  };
  {
    const int pre = (int)(n_my < q.stages ? n_my : q.stages);
    for (int k = 0; k < pre; ++k) issue(k, (gw + k * GW) * q.rps);
  }

  float best = -INFINITY;
  uint32_t brow = kNoRowU;
  int slot = 0;
  uint32_t parity = 0;
  int64_t row0 = gw * q.rps;
  const int64_t ahead = (int64_t)q.stages * row_step;
  for (int64_t k = 0; k < n_my; ++k) {
    mbar_wait(&bars[slot], parity);
    const int64_t left = q.n_rows - row0;
    const int nr = (int)(left < q.rps ? left : q.rps);
    const float* tile = wbuf + (size_t)slot * stage_floats;
    for (int b0 = 0; b0 < nr; b0 += Core::RB)
      Core::batch(tile + (size_t)b0 * q.ld, q.ld, nchunk, g, grp, nr - b0, (uint32_t)(row0 + b0), d0, d1, best, brow);
    __syncwarp();   // every lane's shared-memory reads of this stage are complete
    if (k + q.stages < n_my) issue(slot, row0 + ahead);
    row0 += row_step;
    if (++slot == q.stages) { slot = 0; parity ^= 1u; }
  }
  Core::warp_merge(best, brow);
  if (lane == 0) {
    ScanCand c;
    c.score = best;
    c.row = brow;
    a.cands[gw] = c;
  }
}

}  // namespace bcg
