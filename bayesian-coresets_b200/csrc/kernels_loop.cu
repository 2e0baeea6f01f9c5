// Instantiations + launcher of the persistent greedy-loop kernel (see loop_kernel.cuh).
#include "loop_kernel.cuh"
#include "omp_loop_kernel.cuh"

namespace bcg {

// (CH, LPR, R, J, CH16): J = float64 S-vector elements per control-warp lane (S <= 32 J); CH16 = 16-byte groups per lane of
// the float16 pre-filter pass (0: rows too short for it to pay).  A variant with CH16 > 0 exists with and without the filter.
#define BCG_LOOP_VARIANTS(X) X(1, 32, 8, 4, 0) X(2, 32, 8, 8, 1) X(4, 32, 4, 16, 2) X(1, 16, 8, 2, 0)

bool loop_variant_exists(int ch, int lpr) {
#define X(CH, LPR, R, J, F) if (ch == CH && lpr == LPR) return true;
  BCG_LOOP_VARIANTS(X)
#undef X
  return false;
}

int loop_variant_ch16(int ch, int lpr) {
#define X(CH, LPR, R, J, F) if (ch == CH && lpr == LPR) return F;
  BCG_LOOP_VARIANTS(X)
#undef X
  return 0;
}

// calls FN(kernel) with the instantiation selected by the configuration; `f16`: use the float16 pre-filter variant
template <typename FN>
static cudaError_t with_loop_kernel(const ScanConfig& c, bool f16, FN fn) {
#define X(CH, LPR, R, J, F)                                                            \
  if (c.ch == CH && c.lpr == LPR) {                                                    \
    if (f16 && F > 0) {                                                                \
      if (c.ndir == 2) return fn((const void*)greedy_loop_kernel<CH, 2, LPR, R, J, F>); \
      return fn((const void*)greedy_loop_kernel<CH, 1, LPR, R, J, F>);                 \
    }                                                                                  \
    if (c.ndir == 2) return fn((const void*)greedy_loop_kernel<CH, 2, LPR, R, J, 0>);   \
    return fn((const void*)greedy_loop_kernel<CH, 1, LPR, R, J, 0>);                   \
  }
  BCG_LOOP_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

template <typename FN>
static cudaError_t with_omp_kernel(const ScanConfig& c, bool f16, FN fn) {
#define X(CH, LPR, R, J, F)                                                            \
  if (c.ch == CH && c.lpr == LPR) {                                                    \
    if (f16 && F > 0) return fn((const void*)omp_loop_kernel<CH, LPR, R, F>);          \
    return fn((const void*)omp_loop_kernel<CH, LPR, R, 0>);                            \
  }
  BCG_LOOP_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t loop_set_smem(const ScanConfig& c) {
  for (int f = 0; f < 2; ++f) {
    const cudaError_t e = with_loop_kernel(c, f != 0, [&](const void* k) {
      return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.loop_smem);
    });
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t loop_max_blocks_per_sm(const ScanConfig& c, int* nb) {
  const int threads = (c.wpb + 1) * 32;
  int n0 = 0, n1 = 0;
  cudaError_t e = with_loop_kernel(c, false, [&](const void* k) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, k, threads, c.loop_smem);
  });
  if (e != cudaSuccess) return e;
  e = with_loop_kernel(c, true, [&](const void* k) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k, threads, c.loop_smem);
  });
  *nb = n0 < n1 ? n0 : n1;
  return e;
}

cudaError_t loop_launch(const ScanConfig& c, const LoopArgs& a, cudaStream_t st) {
  void* args[] = {const_cast<LoopArgs*>(&a)};
  const dim3 grid(c.grid), block((c.wpb + 1) * 32);
  return with_loop_kernel(c, a.g.An16 != nullptr, [&](const void* k) {
    return cudaLaunchCooperativeKernel(k, grid, block, args, c.loop_smem, st);
  });
}

// ---- persistent OrthoPursuit kernel (omp_loop_kernel.cuh): same (CH, LPR, R) variants, one direction ----------------
static size_t omp_loop_smem(const ScanConfig& c) {
  const size_t control = (256 + (size_t)(kOmpLoopThreads / 32) * kWideCols) * sizeof(double);
  return c.loop_smem > control ? c.loop_smem : control;
}

cudaError_t omp_loop_set_smem(const ScanConfig& c) {
  for (int f = 0; f < 2; ++f) {
    const cudaError_t e = with_omp_kernel(c, f != 0, [&](const void* k) {
      return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)omp_loop_smem(c));
    });
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t omp_loop_max_blocks_per_sm(const ScanConfig& c, int* nb) {
  int n0 = 0, n1 = 0;
  cudaError_t e = with_omp_kernel(c, false, [&](const void* k) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, k, kOmpLoopThreads, omp_loop_smem(c));
  });
  if (e != cudaSuccess) return e;
  e = with_omp_kernel(c, true, [&](const void* k) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k, kOmpLoopThreads, omp_loop_smem(c));
  });
  *nb = n0 < n1 ? n0 : n1;
  return e;
}

cudaError_t omp_loop_launch(const ScanConfig& c, const LoopArgs& a, NnlsWork* W, int wide, cudaStream_t st) {
  void* args[] = {const_cast<LoopArgs*>(&a), &W, &wide};
  const dim3 grid(c.grid), block(kOmpLoopThreads);
  return with_omp_kernel(c, a.g.An16 != nullptr, [&](const void* k) {
    return cudaLaunchCooperativeKernel(k, grid, block, args, omp_loop_smem(c), st);
  });
}

}  // namespace bcg
