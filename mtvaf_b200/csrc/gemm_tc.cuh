// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma kind::f16, fp32 accumulators in TMEM),
// operands staged by TMA (SWIZZLE_128B) through an mbarrier ring, persistent over output tiles with a
// double-buffered TMEM accumulator so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Replaces the cuBLAS calls behind nn.Linear / torch.matmul on the reference path
// (models/modeling_roberta.py:202,219-220,296,365,379; models/bert_model.py:446-454,458-459;
// probes/probe.py:74) and their autograd dgrad / wgrad.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2..9 = epilogue: two warps per TMEM lane quadrant, each taking half of the tile's
// columns (TMEM -> registers -> fused epilogue -> global).  EPI >= 0 fixes the epilogue at compile time.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/mtvaf_b200.h"
#include "epilogue.cuh"

namespace mtvaf {

constexpr int BM = 128;      // UMMA_M (cta_group::1)
constexpr int BK = 64;       // one 128-byte swizzle atom of bf16 along K
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;   // 16 KB
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024;   // + alignment slack
};

// work item -> (split, m tile, n tile); n fastest so that CTAs running together share the A rows
struct WorkItem {
  int m0, tn, kb0, kb1;
};
__device__ __forceinline__ WorkItem decode_item(int item, int n_tiles_n, int n_tiles_m, int kb_total, int kb_per) {
  WorkItem w;
  w.tn = item % n_tiles_n;
  const int rest = item / n_tiles_n;
  const int tm = rest % n_tiles_m;
  const int sp = rest / n_tiles_m;
  w.m0 = tm * BM;
  w.kb0 = sp * kb_per;
  w.kb1 = min(kb_total, w.kb0 + kb_per);
  return w;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const EpiArgs ep_in, int M, int N, int K, int splits, int kb_per) {
  const EpiArgs ep = resolve_step(ep_in);
  using namespace ptx;
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full = empty_bar + S::kStages;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tiles_m = (M + BM - 1) / BM;
  const int n_tiles_n = (N + BN - 1) / BN;
  const int kb_total = (K + BK - 1) / BK;
  const int n_items = n_tiles_m * n_tiles_n * splits;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  constexpr int kTmemCols = 2 * BN;   // 256 / 512: powers of two
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const WorkItem w = decode_item(item, n_tiles_n, n_tiles_m, kb_total, kb_per);
        const int n0 = w.tn * BN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = tiles + stage * S::kStageBytes;
          uint8_t* sB = sA + S::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          const int k0 = kb * BK;
          if (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[stage], k0, w.m0);            // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)                             // boxes {64 m, 64 k}
              tma_load_2d(sA + j * 8192, &tmA, &full_bar[stage], w.m0 + j * 64, k0);
          }
          if (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], k0, n0);              // box {64 k, BN n}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)                             // boxes {64 n, 64 k}
              tma_load_2d(sB + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, k0);
          }
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const WorkItem w = decode_item(item, n_tiles_n, n_tiles_m, kb_total, kb_per);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(tiles + stage * S::kStageBytes);
          const uint32_t sB = sA + S::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = A_MN ? make_smem_desc_sw128(sA + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sA + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(sB + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(sB + k * 32, 16, 1024);
            umma_f16_ss(d_tmem, da, db, idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);          // frees this smem stage when the MMAs retire
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);              // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;              // which half of the tile's columns
    constexpr int kChunks = BN / 64;               // 32-column chunks per half
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const WorkItem w = decode_item(item, n_tiles_n, n_tiles_m, kb_total, kb_per);
      const int n0 = w.tn * BN + half * (BN / 2);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = w.m0 + quad * 32 + lane;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * (BN / 2);
      float rowacc = 0.f;
#pragma unroll 2
      for (int c = 0; c < kChunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_addr + c * 32, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (col0 < N) epilogue_row32<EPI>(ep, r, row, col0, M, N, rowacc);
      }
      epilogue_row_finish<EPI>(ep, row, M, rowacc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// 2D bf16 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, row stride `ld` elements
int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer);

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch_gemm_tc(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const EpiArgs& ep,
                   int splits, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_tmap_bf16_2d(&tmA, A, K, M, lda, BK, BM);
  else       rc = make_tmap_bf16_2d(&tmA, A, M, K, lda, 64, BK);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_bf16_2d(&tmB, B, K, N, ldb, BK, BN);
  else       rc = make_tmap_bf16_2d(&tmB, B, N, K, ldb, 64, BK);
  if (rc) return rc;

  const int kb_total = (K + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  const int kb_per = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per - 1) / kb_per;
  const int n_items = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * splits;
  const int grid = n_items < sm_count() ? n_items : sm_count();

  auto kern = gemm_bf16_tc_kernel<BN, A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  kern<<<grid, kGemmThreads, S::kTotal, stream>>>(tmA, tmB, ep, M, N, K, splits, kb_per);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

// per-layout dispatchers (one translation unit each, so nvcc compiles them in parallel)
int gemm_tc_kk(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const EpiArgs& ep,
               int splits, cudaStream_t stream);
int gemm_tc_kmn(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const EpiArgs& ep,
                int splits, cudaStream_t stream);
int gemm_tc_mnmn(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const EpiArgs& ep,
                 int splits, cudaStream_t stream);

#define MTVAF_GEMM_CASE(MODE_, AMN_, BMN_)                                                              \
  case MODE_:                                                                                           \
    return narrow ? launch_gemm_tc<128, AMN_, BMN_, MODE_>(A, lda, B, ldb, M, N, K, ep, splits, stream) \
                  : launch_gemm_tc<256, AMN_, BMN_, MODE_>(A, lda, B, ldb, M, N, K, ep, splits, stream)

}  // namespace mtvaf
