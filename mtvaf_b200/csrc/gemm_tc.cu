// Host side of the tcgen05 GEMM: TMA tensor-map construction and the C-ABI entry point.
#include "gemm_tc2.cuh"

namespace mtvaf {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled fn = get_encode_fn();
  MTVAF_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  MTVAF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  MTVAF_REQUIRE((ld * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes (ld=%llu)",
                (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MTVAF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu box=%ux%u",
                (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner,
                box_outer);
  return 0;
}

}  // namespace mtvaf

using namespace mtvaf;

static int g_gemm_impl = 0;
namespace mtvaf { int gemm_impl_override() { return g_gemm_impl; } }
extern "C" int mtvaf_set_gemm_impl(int impl) {
  MTVAF_REQUIRE(impl == 0 || impl == 1, "gemm impl must be 0 (auto: CTA pairs) or 1 (single-CTA tiles)");
  g_gemm_impl = impl;
  return 0;
}

extern "C" int mtvaf_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                               int b_mn_major, int M, int N, int K, const MtvafEpilogue* epi, int splits,
                               void* stream) {
  MTVAF_REQUIRE(A && B && epi, "mtvaf_gemm_bf16: null argument");
  MTVAF_REQUIRE(M > 0 && N > 0 && K > 0, "mtvaf_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  EpiArgs ep;
  int rc = make_epi_args(*epi, MTVAF_BF16, M, N, splits, &ep);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // optional column sums of the output: inside the TMA-staged epilogue of the CTA-pair kernel, else a pass over `out`
  const bool staged_mode = ep.mode == MTVAF_EPI_STORE || ep.mode == MTVAF_EPI_GELU || ep.mode == MTVAF_EPI_TANH ||
                           ep.mode == MTVAF_EPI_RESID || ep.mode == MTVAF_EPI_MUL_DGELU || ep.mode == MTVAF_EPI_MUL_DTANH ||
                           ep.mode == MTVAF_EPI_GELU_GRAD || ep.mode == MTVAF_EPI_MUL_AUX;
  const bool fused = ep.colsum && ep.staged && staged_mode && M >= 256 && gemm_impl_override() == 0;
  float* post = fused ? nullptr : ep.colsum;
  if (!fused) ep.colsum = nullptr;
  if (!a_mn_major && !b_mn_major) rc = gemm_tc_kk(A, lda, B, ldb, M, N, K, ep, splits, st);
  else if (!a_mn_major && b_mn_major) rc = gemm_tc_kmn(A, lda, B, ldb, M, N, K, ep, splits, st);
  else if (a_mn_major && b_mn_major) rc = gemm_tc_mnmn(A, lda, B, ldb, M, N, K, ep, splits, st);
  else {
    MTVAF_REQUIRE(false, "mtvaf_gemm_bf16: A MN-major with B K-major is not used by the path");
  }
  if (rc == 0 && post) rc = mtvaf_colsum(epi->out, epi->ldo, epi->out_dtype, M, N, post, stream);
  return rc;
}
