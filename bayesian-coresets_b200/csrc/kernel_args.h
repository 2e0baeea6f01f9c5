// Launch-argument structs and the internal launcher interface between the translation units of
// libbcg_b200.so (bcg_api.cu, kernels_scan.cu, kernels_loop.cu).  Host-compilable.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "bcg_state.h"

namespace bcg {

// geometry of the streamed matrix and of the per-warp TMA ring
struct ScanGeom {
  const float* An;      // n_rows x ld, unit rows
  int64_t n_rows;
  int32_t ld;           // floats per row (multiple of 4)
  int32_t rps;          // rows per stage (multiple of the batch size R * 32/LPR)
  int32_t stages;
  int32_t evict_first;  // L2 policy of the streaming loads
  // float16 pre-filter (persistent kernels only; null = off): the ring streams this copy instead of An
  const uint16_t* An16; // n_rows x ld16 halves (IEEE binary16, round to nearest of An; padding zero)
  int32_t ld16;         // halves per row (multiple of 8)
  int32_t rps16;        // rows per ring stage in the float16 pass
  float eps16;          // filter_eps_unit(S)
};

struct ScanArgs {
  ScanGeom g;
  const float* dir;       // NDIR x ld
  ScanCand* cands;        // gridDim.x * warps_per_block
  const int32_t* skip0;   // scan is skipped when *skip0 or *skip1 is non-zero (halted / select failed)
  const int32_t* skip1;
  float* lost;            // gridDim.x * warps_per_block: best score each warp saw but did not publish
  unsigned int* done;     // CTA completion counter (zero between launches)
  int32_t* need_exact;    // raised by the last CTA when the candidate set may miss the float64 arg-max
  const int32_t* force_exact;
  float* top;             // summary of this scan (SolverState::scan_top / scan_top_row / scan_cnt)
  uint32_t* top_row;
  int32_t* top_cnt;
};

struct LoopCtl {
  unsigned int arrive;   // number of CTA arrivals so far (monotonic within a launch)
  unsigned int go;       // iterations published so far
  unsigned int stop;     // set (before go) when the loop ends early
  unsigned int pad;
};

struct LoopArgs {
  SolverState* st;
  LoopCtl* ctl;
  ScanCand* cta_cands;   // 2 per CTA: best and runner-up of the CTA's warps
  float* cta_lost;       // 1 per CTA: best score the CTA saw but did not publish
  int32_t cont;          // continuation of the same build() call after an exact-selection stop (keeps the retry flag)
  int32_t use_pre;       // the first iteration takes its local winner from st->exact_cands (exact_scan_kernel)
  ScanGeom g;
  int32_t itrs;
  int32_t wpb;           // scan warps per CTA (blockDim.x = (wpb + 1) * 32)
  unsigned int* claims;  // itrs zero-initialised counters: dynamically claimed chunks per iteration
  float static_frac;     // share of the chunks that is statically assigned (rest: dynamic tail)
  int* filt_L;           // itrs words preset to 0x80808080: GPU-wide lower bound of the float32 maximum per iteration (float16
                         // pre-filter; monotone int image of the float, raised with atomicMax)
  // optional device timestamps (globaltimer ns), null when tracing is off:
  //   trace[it*8 + 0] control: grid arrived      trace[it*8 + 1] control: next direction published
  //   trace[it*8 + 2] CTA 0 warp 0: go observed   trace[it*8 + 3] CTA 0 warp 0: its scan finished
  //   trace[it*8 + 4..7] control: candidates reduced / winning row fetched / line search done / committed
  unsigned long long* trace;
};

// template selection + launch geometry, chosen on the host from the row length
struct ScanConfig {
  int ch, ndir, lpr, r;      // template parameters of the scan core
  int rb;                    // rows per batch = r * 32 / lpr
  int rps, stages, wpb, grid, evict_first;
  int ch16, rps16;           // float16 pre-filter: 16-byte groups per lane (0 = not available for this row length)
  size_t smem;               // dynamic shared memory of scan_kernel
  size_t loop_smem;          // dynamic shared memory of greedy_loop_kernel
};

// kernels_scan.cu
bool scan_variant_exists(int ch, int lpr);
int scan_variant_r(int ch, int lpr);
cudaError_t scan_set_smem(const ScanConfig& c);
cudaError_t scan_launch(const ScanConfig& c, const ScanArgs& a, cudaStream_t st);
cudaError_t exact_scan_launch(int grid, SolverState* st, int force, cudaStream_t s);
// kernels_loop.cu
bool loop_variant_exists(int ch, int lpr);
int loop_variant_ch16(int ch, int lpr);   // 16-byte groups per lane of the float16 pre-filter (0: none)
cudaError_t loop_set_smem(const ScanConfig& c);
cudaError_t loop_max_blocks_per_sm(const ScanConfig& c, int* nb);
cudaError_t loop_launch(const ScanConfig& c, const LoopArgs& a, cudaStream_t st);
struct NnlsWork;
cudaError_t omp_loop_set_smem(const ScanConfig& c);
cudaError_t omp_loop_max_blocks_per_sm(const ScanConfig& c, int* nb);
cudaError_t omp_loop_launch(const ScanConfig& c, const LoopArgs& a, NnlsWork* W, int wide, cudaStream_t st);

}  // namespace bcg
