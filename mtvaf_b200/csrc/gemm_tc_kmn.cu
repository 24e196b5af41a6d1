// tcgen05 GEMM instantiations: A K-major, B MN-major (dgrad dX = dY W and fused backward epilogues)
#include "gemm_tc2.cuh"
namespace mtvaf {
int gemm_tc_kmn(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const EpiArgs& ep,
                int splits, cudaStream_t stream) {
  const bool narrow = (N <= 128);
  const bool pair = (M >= 256) && gemm_impl_override() == 0;
  switch (ep.mode) {
    MTVAF_GEMM_CASE2(MTVAF_EPI_STORE, false, true);
    MTVAF_GEMM_CASE2(MTVAF_EPI_RESID, false, true);
    MTVAF_GEMM_CASE2(MTVAF_EPI_MUL_DGELU, false, true);
    MTVAF_GEMM_CASE2(MTVAF_EPI_MUL_DTANH, false, true);
    MTVAF_GEMM_CASE2(MTVAF_EPI_MUL_AUX, false, true);
    default:
      return narrow ? launch_gemm_tc<128, false, true, -1>(A, lda, B, ldb, M, N, K, ep, splits, stream)
                    : launch_gemm_tc<256, false, true, -1>(A, lda, B, ldb, M, N, K, ep, splits, stream);
  }
}
}  // namespace mtvaf
