// tcgen05 GEMM instantiations: A MN-major, B MN-major (wgrad dW += dY^T X, split-K fp32 atomics)
#include "gemm_tc2.cuh"
namespace mtvaf {
int gemm_tc_mnmn(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const EpiArgs& ep,
                 int splits, cudaStream_t stream) {
  const bool narrow = (N <= 128);
  const bool pair = (M >= 256) && gemm_impl_override() == 0;
  switch (ep.mode) {
    MTVAF_GEMM_CASE2(MTVAF_EPI_ATOMIC_F32, true, true);
    default:
      return narrow ? launch_gemm_tc<128, true, true, -1>(A, lda, B, ldb, M, N, K, ep, splits, stream)
                    : launch_gemm_tc<256, true, true, -1>(A, lda, B, ldb, M, N, K, ep, splits, stream);
  }
}
}  // namespace mtvaf
