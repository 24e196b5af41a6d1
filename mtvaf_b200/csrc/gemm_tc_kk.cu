// tcgen05 GEMM instantiations: A K-major, B K-major (forward Y = X W^T and fused-epilogue variants)
#include "gemm_tc2.cuh"
namespace mtvaf {
int gemm_tc_kk(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const EpiArgs& ep,
               int splits, cudaStream_t stream) {
  const bool narrow = (N <= 128);
  const bool pair = (M >= 256) && gemm_impl_override() == 0;
  switch (ep.mode) {
    MTVAF_GEMM_CASE2(MTVAF_EPI_STORE, false, false);
    MTVAF_GEMM_CASE2(MTVAF_EPI_GELU, false, false);
    MTVAF_GEMM_CASE2(MTVAF_EPI_GELU_GRAD, false, false);
    MTVAF_GEMM_CASE2(MTVAF_EPI_TANH, false, false);
    MTVAF_GEMM_CASE2(MTVAF_EPI_RESID, false, false);
    MTVAF_GEMM_CASE2(MTVAF_EPI_SQNORM, false, false);
    MTVAF_GEMM_CASE2(MTVAF_EPI_ROWSCALE, false, false);
    default:
      return narrow ? launch_gemm_tc<128, false, false, -1>(A, lda, B, ldb, M, N, K, ep, splits, stream)
                    : launch_gemm_tc<256, false, false, -1>(A, lda, B, ldb, M, N, K, ep, splits, stream);
  }
}
}  // namespace mtvaf
