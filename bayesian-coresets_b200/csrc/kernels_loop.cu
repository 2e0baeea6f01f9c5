// Instantiations + launcher of the persistent greedy-loop kernel (see loop_kernel.cuh).
#include "loop_kernel.cuh"
#include "omp_loop_kernel.cuh"

namespace bcg {

// (CH, LPR, R, J): J = float64 S-vector elements per control-warp lane (S <= 32 J)
#define BCG_LOOP_VARIANTS(X) X(1, 32, 8, 4) X(2, 32, 8, 8) X(4, 32, 4, 16) X(1, 16, 8, 2)

bool loop_variant_exists(int ch, int lpr) {
#define X(CH, LPR, R, J) if (ch == CH && lpr == LPR) return true;
  BCG_LOOP_VARIANTS(X)
#undef X
  return false;
}

cudaError_t loop_set_smem(const ScanConfig& c) {
#define X(CH, LPR, R, J)                                                                                       \
  if (c.ch == CH && c.lpr == LPR) {                                                                            \
    if (c.ndir == 2)                                                                                           \
      return cudaFuncSetAttribute(greedy_loop_kernel<CH, 2, LPR, R, J>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)c.loop_smem);                                                           \
    return cudaFuncSetAttribute(greedy_loop_kernel<CH, 1, LPR, R, J>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                (int)c.loop_smem);                                                             \
  }
  BCG_LOOP_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t loop_max_blocks_per_sm(const ScanConfig& c, int* nb) {
  const int threads = (c.wpb + 1) * 32;
#define X(CH, LPR, R, J)                                                                                          \
  if (c.ch == CH && c.lpr == LPR) {                                                                               \
    if (c.ndir == 2)                                                                                              \
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, greedy_loop_kernel<CH, 2, LPR, R, J>, threads, c.loop_smem); \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, greedy_loop_kernel<CH, 1, LPR, R, J>, threads, c.loop_smem);   \
  }
  BCG_LOOP_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t loop_launch(const ScanConfig& c, const LoopArgs& a, cudaStream_t st) {
  void* args[] = {const_cast<LoopArgs*>(&a)};
  const dim3 grid(c.grid), block((c.wpb + 1) * 32);
#define X(CH, LPR, R, J)                                                                                         \
  if (c.ch == CH && c.lpr == LPR) {                                                                              \
    if (c.ndir == 2)                                                                                             \
      return cudaLaunchCooperativeKernel((const void*)greedy_loop_kernel<CH, 2, LPR, R, J>, grid, block, args,  \
                                         c.loop_smem, st);                                                       \
    return cudaLaunchCooperativeKernel((const void*)greedy_loop_kernel<CH, 1, LPR, R, J>, grid, block, args,    \
                                       c.loop_smem, st);                                                         \
  }
  BCG_LOOP_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

// ---- persistent OrthoPursuit kernel (omp_loop_kernel.cuh): same (CH, LPR, R) variants, one direction ----------------
static size_t omp_loop_smem(const ScanConfig& c) {
  const size_t control = (256 + (size_t)(kOmpLoopThreads / 32) * kWideCols) * sizeof(double);
  return c.loop_smem > control ? c.loop_smem : control;
}

cudaError_t omp_loop_set_smem(const ScanConfig& c) {
#define X(CH, LPR, R, J)                                                                                        \
  if (c.ch == CH && c.lpr == LPR)                                                                               \
    return cudaFuncSetAttribute(omp_loop_kernel<CH, LPR, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)omp_loop_smem(c));
  BCG_LOOP_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t omp_loop_max_blocks_per_sm(const ScanConfig& c, int* nb) {
#define X(CH, LPR, R, J)                                                                                        \
  if (c.ch == CH && c.lpr == LPR)                                                                               \
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, omp_loop_kernel<CH, LPR, R>, kOmpLoopThreads, omp_loop_smem(c));
  BCG_LOOP_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t omp_loop_launch(const ScanConfig& c, const LoopArgs& a, NnlsWork* W, int wide, cudaStream_t st) {
  void* args[] = {const_cast<LoopArgs*>(&a), &W, &wide};
  const dim3 grid(c.grid), block(kOmpLoopThreads);
#define X(CH, LPR, R, J)                                                                                        \
  if (c.ch == CH && c.lpr == LPR)                                                                               \
    return cudaLaunchCooperativeKernel((const void*)omp_loop_kernel<CH, LPR, R>, grid, block, args, omp_loop_smem(c), st);
  BCG_LOOP_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace bcg
