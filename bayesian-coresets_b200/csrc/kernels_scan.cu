// Instantiations + launcher of the stand-alone scan kernel (see scan_kernel.cuh).
#include "scan_kernel.cuh"

namespace bcg {

// (CH, LPR, R) variants: LPR = 32 with 1/2/4/8 float4 chunks per lane (S <= 128/256/512/1024),
// and narrow rows (S <= 64/32/16/8/4) with 16/8/4/2/1 lanes per row
#define BCG_SCAN_VARIANTS(X) X(1, 32, 8) X(2, 32, 8) X(4, 32, 4) X(8, 32, 2) X(1, 16, 8) X(1, 8, 8) X(1, 4, 4) X(1, 2, 2) X(1, 1, 8)

bool scan_variant_exists(int ch, int lpr) {
#define X(CH, LPR, R) if (ch == CH && lpr == LPR) return true;
  BCG_SCAN_VARIANTS(X)
#undef X
  return false;
}

int scan_variant_r(int ch, int lpr) {
#define X(CH, LPR, R) if (ch == CH && lpr == LPR) return R;
  BCG_SCAN_VARIANTS(X)
#undef X
  return 0;
}

cudaError_t scan_set_smem(const ScanConfig& c) {
#define X(CH, LPR, R)                                                                                              \
  if (c.ch == CH && c.lpr == LPR) {                                                                                \
    if (c.ndir == 2)                                                                                               \
      return cudaFuncSetAttribute(scan_kernel<CH, 2, LPR, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem); \
    return cudaFuncSetAttribute(scan_kernel<CH, 1, LPR, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);   \
  }
  BCG_SCAN_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t scan_launch(const ScanConfig& c, const ScanArgs& a, cudaStream_t st) {
#define X(CH, LPR, R)                                                                   \
  if (c.ch == CH && c.lpr == LPR) {                                                     \
    if (c.ndir == 2)                                                                    \
      scan_kernel<CH, 2, LPR, R><<<c.grid, c.wpb * 32, c.smem, st>>>(a);                \
    else                                                                                \
      scan_kernel<CH, 1, LPR, R><<<c.grid, c.wpb * 32, c.smem, st>>>(a);                \
    return cudaGetLastError();                                                          \
  }
  BCG_SCAN_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t exact_scan_launch(int grid, SolverState* st, int force, cudaStream_t s) {
  exact_scan_kernel<<<grid, 512, 0, s>>>(st, force);
  return cudaGetLastError();
}

}  // namespace bcg
