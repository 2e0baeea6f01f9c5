// K3 on the float64 tensor cores: the MATERIALISING projection (unit float32 rows + float64 norms + column sums)
// with the z . theta contraction issued as DMMA (mma.sync.m8n8k4.f64).
//
// Same reference lines and the same float64 arithmetic as project_fast_kernel.cuh (projector.py:19-21 around
// model_lr.py:25-32 / model_gaussian.py:4-10 / model_poiss.py:25-38, giga.py:10-13, hilbert.py:24).
//
// Why (ncu of project_fast_kernel<8, LR> at d = 10, S = 512, profiles/r01b_project_fast_lr_chunk_S512_full.csv and the
// l1tex / fp64 metrics of profiles/r02_*): with one warp per row every lane needs ITS OWN 16 columns of the sample
// tile for every k -- 80 LDS.128 per lane and row, 320 of the ~420 L1 wavefronts a row costs -- and the L1 data pipe, not
// the float64 pipe (36 % busy) or HBM (12 %), bounded the kernel at 24 ms for N = 1e7.  A DMMA B-fragment is shared by
// the 8 rows of the tile, so the sample tile is read once per EIGHT rows: 6 wavefronts per row instead of 320.
//
// Decomposition.  CTA = 8 warps over a block of BM = 8 MI WR rows x S columns, S in {64, 128, 256, 512}: warp (wr, wc)
// owns rows wr*8MI .. and the 64-column strip wc (8 n-tiles), i.e. MI x 8 DMMA tiles = MI*16 accumulators per lane;
// lane (g, tq) of a tile holds row g, columns 2 tq, 2 tq + 1.  The row-wide reductions of the projection (mean over the S
// samples, row norm) cross the WC = S/64 strip warps: quad shuffle -> shared memory -> block barrier, two barriers per
// row block; with 2-3 CTAs per SM one CTA's barrier wait is another CTA's link evaluation.
//   MI = 1: d <= 16 -- the whole (zero-padded) sample tile stays resident in shared memory (66 KB at S = 512);
//           A fragments come straight from global memory (8 rows x 32 B per load, L1 hits for 7 of the 8 warps)
//   MI = 4: any d   -- the sample tile streams through a double-buffered 16 x S shared-memory tile per k step, one
//           barrier per k tile (as project_sum_mma_kernel); 32 rows per block amortise each tile over 4 row tiles.
//           This is the selection pass of SparseVI at d = 200 (sparsevi.py:30,51): a genuine dense contraction.
// Column sums accumulate in registers per lane and are reduced lane -> warp -> CTA in a fixed order at the end
// (bit-reproducible b).  Bytes per row: 8 d_in read + 4 S + 8 written; bound: float64 pipe (link evaluation).
#pragma once
#include "project_fast_kernel.cuh"
#include "project_sum_mma_kernel.cuh"

namespace bcg {

constexpr int kPjThreads = 256;
constexpr int kPjKT = 16;                         // k rows per shared-memory sample tile

inline size_t project_mma_smem(int S, int mi) {
  const size_t ld = (size_t)S + 4;
  const int WC = S / 64, WR = 8 / WC;
  const size_t tile = (size_t)kPjKT * ld * sizeof(double);
  const size_t red = (size_t)2 * (8 * mi * WR) * WC * sizeof(double);
  const size_t comb = (size_t)WR * S * sizeof(double) + 64;
  return (mi == 1 ? tile : 2 * tile) + red + comb;
}

template <int MODEL, int MI>
__global__ void __launch_bounds__(kPjThreads, MI == 1 ? 2 : 1) project_mma_kernel(const ProjectArgs a) {
  extern __shared__ __align__(16) double pj_smem[];
  const int S = a.S, d = a.d;
  const int ld = S + 4;                            // padded sample-tile row: conflict-free B-fragment loads
  const int WC = S >> 6, WR = 8 / WC;
  const int BM = 8 * MI * WR;
  constexpr int NBUF = MI == 1 ? 1 : 2;
  double* ths = pj_smem;                                            // [NBUF][kPjKT][ld]
  double* red = ths + (size_t)NBUF * kPjKT * ld;                    // [2][BM][WC]
  double* comb = red + (size_t)2 * BM * WC;                         // [WR][S] + 8 (end-of-kernel column-sum combine)
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int wc = warp % WC, wr = warp / WC;
  const int col0 = wc * 64;
  const double* __restrict__ tab = a.sp_tab;
  const int nk = (d + kPjKT - 1) / kPjKT;
  const int64_t nblocks = (a.n + BM - 1) / BM;

  auto load_tile = [&](int buf, int k0) {          // sample rows k0 .. k0+15 (zero beyond d) -> shared
    double* dst = ths + (size_t)buf * kPjKT * ld;
#pragma unroll 4
    for (int k = 0; k < kPjKT; ++k) {
      const bool in = k0 + k < d;
      for (int c = t; c < S; c += kPjThreads) dst[(size_t)k * ld + c] = in ? __ldg(a.theta + (size_t)(k0 + k) * S + c) : 0.;
    }
  };
  if (MI == 1) {
    load_tile(0, 0);
    __syncthreads();
  }

  double cs[8][2];
#pragma unroll
  for (int ni = 0; ni < 8; ++ni) cs[ni][0] = cs[ni][1] = 0.;
  double normsum = 0.;

  // MI == 1: the A fragments (and y) of the NEXT row block are fetched while the current block's links are evaluated
  double afn[kPjKT / 4], yn = 0.;
  auto fetch_block = [&](int64_t rb) {
#pragma unroll
    for (int q = 0; q < kPjKT / 4; ++q) afn[q] = 0.;
    yn = 0.;
    const int64_t r = rb * BM + wr * 8 + g;
    if (rb < nblocks && r < a.n) {
      const int64_t z = a.rowidx ? a.rowidx[r] : r;
#pragma unroll
      for (int q = 0; q < kPjKT / 4; ++q)
        if (q * 4 + tq < d) afn[q] = __ldg(a.Z + z * a.zld + q * 4 + tq);
      if (MODEL == MODEL_POISSON) yn = __ldg(a.Z + z * a.zld + d);
    }
  };
  if (MI == 1) fetch_block(blockIdx.x);

  for (int64_t rb = blockIdx.x; rb < nblocks; rb += gridDim.x) {
    const int64_t row0 = rb * BM;
    int64_t zr[MI];
    bool live[MI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
      const int64_t r = row0 + wr * (8 * MI) + mi * 8 + g;
      live[mi] = r < a.n;
      zr[mi] = (live[mi] && MI > 1) ? (a.rowidx ? a.rowidx[r] : r) : 0;
    }
    double acc[MI][8][2];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      double c0 = 0., c1 = 0.;
      if (MODEL == MODEL_LINEAR && a.coff) {
        c0 = __ldg(a.coff + col0 + ni * 8 + tq * 2);
        c1 = __ldg(a.coff + col0 + ni * 8 + tq * 2 + 1);
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) { acc[mi][ni][0] = c0; acc[mi][ni][1] = c1; }
    }
    // ---- contraction ---------------------------------------------------------------------------
    double ycur = 0.;
    double afc[kPjKT / 4];                                           // (MI == 1) this block's A fragments, kept for the tail path
    auto contract_resident = [&]() {
#pragma unroll
      for (int kk = 0; kk < kPjKT; kk += 4) {
        if (kk < d) {                                                // block-uniform
          const double* brow = ths + (size_t)(kk + tq) * ld + col0 + g;
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) dmma_m8n8k4(acc[0][ni][0], acc[0][ni][1], afc[kk / 4], brow[ni * 8]);
        }
      }
    };
    if (MI == 1) {
#pragma unroll
      for (int q = 0; q < kPjKT / 4; ++q) afc[q] = afn[q];
      contract_resident();
      ycur = yn;
      fetch_block(rb + gridDim.x);                                   // in flight during this block's epilogue
    } else {
      __syncthreads();                                               // previous block's tiles are no longer read
      load_tile(0, 0);
      __syncthreads();
      double afq[MI];                                                // A fragments, fetched one k step ahead
      auto fetch_a = [&](int k) {
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) afq[mi] = (live[mi] && k < d) ? __ldg(a.Z + zr[mi] * a.zld + k) : 0.;
      };
      fetch_a(tq);
      for (int kt = 0; kt < nk; ++kt) {
        const int k0 = kt * kPjKT;
        if (kt + 1 < nk) load_tile((kt + 1) & 1, k0 + kPjKT);        // the other buffer: last read before the previous barrier
        const double* tile = ths + (size_t)(kt & 1) * kPjKT * ld;
#pragma unroll
        for (int kk = 0; kk < kPjKT; kk += 4) {
          if (k0 + kk < d) {
            double af[MI];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) af[mi] = afq[mi];
            fetch_a(k0 + kk + 4 + tq);
            const double* brow = tile + (size_t)(kk + tq) * ld + col0 + g;
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
              const double bf = brow[ni * 8];
#pragma unroll
              for (int mi = 0; mi < MI; ++mi) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf);
            }
          }
        }
        __syncthreads();
      }
    }
    // ---- link + row mean (projector.py:20-21) ----------------------------------------------------
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
      const double y = MODEL != MODEL_POISSON ? 0. : (MI == 1 ? ycur : (live[mi] ? __ldg(a.Z + zr[mi] * a.zld + d) : 0.));
      double sum = 0.;
      if (MI == 1 && MODEL != MODEL_LINEAR) {
        // branch-free links: 16 independent evaluations the compiler can interleave (softplus_table.h); when any lane of
        // the warp met |lin| > 37 (rare) the tile is contracted again and evaluated with the branching links
        bool tail = false;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double lin = acc[mi][ni][e];
            tail |= link_needs_tail(lin);
            acc[mi][ni][e] = MODEL == MODEL_LR ? lr_link_nb(tab, lin) : poisson_link_nb(tab, lin, y);
          }
        if (__any_sync(0xffffffffu, tail)) {
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.;
          contract_resident();
#pragma unroll
          for (int ni = 0; ni < 8; ++ni) {
            acc[mi][ni][0] = fast_link<MODEL>(tab, acc[mi][ni][0], y);
            acc[mi][ni][1] = fast_link<MODEL>(tab, acc[mi][ni][1], y);
          }
        }
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) sum += acc[mi][ni][0] + acc[mi][ni][1];
      } else {
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
          acc[mi][ni][0] = fast_link<MODEL>(tab, acc[mi][ni][0], y);
          acc[mi][ni][1] = fast_link<MODEL>(tab, acc[mi][ni][1], y);
          sum += acc[mi][ni][0] + acc[mi][ni][1];
        }
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (tq == 0) red[(size_t)wc * BM + wr * (8 * MI) + mi * 8 + g] = sum;      // [strip][row]: conflict-free
    }
    __syncthreads();
    double* red1 = red + (size_t)BM * WC;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
      const double* rr = red + wr * (8 * MI) + mi * 8 + g;
      double tot = 0.;
      for (int w = 0; w < WC; ++w) tot += rr[(size_t)w * BM];
      const double mean = tot * (1. / (double)S);                    // S is a power of two: exact reciprocal
      double ss = 0.;
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        acc[mi][ni][0] -= mean; acc[mi][ni][1] -= mean;
        ss = fma(acc[mi][ni][0], acc[mi][ni][0], ss);
        ss = fma(acc[mi][ni][1], acc[mi][ni][1], ss);
        if (live[mi]) { cs[ni][0] += acc[mi][ni][0]; cs[ni][1] += acc[mi][ni][1]; }
      }
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      if (tq == 0) red1[(size_t)wc * BM + wr * (8 * MI) + mi * 8 + g] = ss;
    }
    __syncthreads();
    // ---- norm, unit float32 row (giga.py:10-13) ----------------------------------------------------
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
      const double* rr = red1 + wr * (8 * MI) + mi * 8 + g;
      double tot = 0.;
      for (int w = 0; w < WC; ++w) tot += rr[(size_t)w * BM];
      const double norm = sqrt(tot);
      const double inv = norm > 0. ? 1. / norm : 0.;
      if (live[mi]) {
        const int64_t r = row0 + wr * (8 * MI) + mi * 8 + g;
        float2* out = reinterpret_cast<float2*>(a.An + (size_t)r * S + col0) + tq;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
          out[ni * 4] = make_float2((float)(acc[mi][ni][0] * inv), (float)(acc[mi][ni][1] * inv));
        if (wc == 0 && tq == 0) {
          a.norms[r] = norm;
          normsum += norm;
          if (norm == 0.) atomicAdd(a.zero_rows, 1ull);
        }
      }
    }
  }

  // ---- column sums: lanes (over g) -> warp rows -> CTA, fixed order ---------------------------------
#pragma unroll
  for (int ni = 0; ni < 8; ++ni)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = cs[ni][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      cs[ni][e] = v;
    }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) normsum += __shfl_xor_sync(0xffffffffu, normsum, off);
  __syncthreads();
  if (g == 0) {
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
      comb[(size_t)wr * S + col0 + ni * 8 + tq * 2] = cs[ni][0];
      comb[(size_t)wr * S + col0 + ni * 8 + tq * 2 + 1] = cs[ni][1];
    }
  }
  double* nsum = comb + (size_t)WR * S;
  if (lane == 0) nsum[warp] = normsum;
  __syncthreads();
  for (int c = t; c < S; c += kPjThreads) {
    double v = 0.;
    for (int w = 0; w < WR; ++w) v += comb[(size_t)w * S + c];
    a.partial[(size_t)blockIdx.x * (S + 1) + c] = v;
  }
  if (t == 0) {
    double v = 0.;
    for (int w = 0; w < 8; ++w) v += nsum[w];
    a.partial[(size_t)blockIdx.x * (S + 1) + S] = v;
  }
}

}  // namespace bcg
