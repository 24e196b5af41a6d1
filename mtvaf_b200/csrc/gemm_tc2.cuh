// bf16 GEMM on tcgen05 with CTA PAIRS (cta_group::2): two CTAs of a 2-CTA cluster (one TPC) cooperate on a
// 256 x BN output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows) per
// k-block, so shared-memory traffic per SM is 2/3 of the single-CTA kernel and the MMA runs at M = 256;
// the leader CTA's elected thread issues `tcgen05.mma.cta_group::2`, each CTA's TMEM receives the 128
// accumulator rows it owns, and each CTA runs its own epilogue warps.
//
// Pipelines (all mbarriers):
//   full[s]   (leader CTA only)  : leader's producer arrives with expect_tx = bytes of BOTH CTAs; both
//                                   CTAs' TMA loads (cp.async.bulk.tensor .cta_group::2) complete_tx on it
//   empty[s]  (one per CTA)      : tcgen05.commit.cta_group::2 multicast to both CTAs frees stage s in both
//   tmem_full[a]  (one per CTA)  : commit multicast -> both CTAs' epilogue warps
//   tmem_empty[a] (leader only)  : 2 x kEpiWarps arrivals (peer epilogue warps arrive remotely)
// Same operand layouts / epilogues / split-K as gemm_tc.cuh (which remains for small problems).
//
// Epilogue: SIXTEEN warps per CTA (576 threads).  The fused epilogues are latency-bound per warp (TMEM load -> math ->
// staging write -> TMA store; issue slots were ~50 % used with two epilogue warps per scheduler, and the K = 768 tiles
// were EPILOGUE-bound: tensor pipe 47-58 %, profiles/r1_ncu_hot_v15.md), so the fix is thread-level parallelism, not
// fewer instructions: warps work in PAIRS on one [32 rows x 64 cols] box -- each warp takes 32 of the 64 columns, in
// two 16-column steps to stay under 112 registers -- and share the pair's staging / aux boxes (shared memory per CTA
// is unchanged), synchronised by one named barrier per pair.
#pragma once
#include "gemm_tc.cuh"
#include "epilogue_staged.cuh"

namespace mtvaf {

constexpr int kEpiWarps2 = 16;                       // epilogue warps per CTA of the pair kernel
constexpr int kEpiPairs = kEpiWarps2 / 2;            // warp pairs = owners of the staging / aux boxes
constexpr int kGemmThreads2 = 64 + kEpiWarps2 * 32;  // + TMA producer warp + MMA issuer warp

namespace ptx {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Signal without a memory fence.  For "this TMEM accumulator has been read": the reads are already complete
// (tcgen05.wait::ld) and ordered by tcgen05.fence::before_thread_sync; nothing written to memory has to be visible
// to the MMA issuer.  The .release.cluster form costs MEMBAR.ALL.GPU + ERRBAR per arrival -- 25 % of all warp
// samples of the fused-epilogue GEMMs (profiles/r1_ncu_hot_v7.md).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// TMA load into THIS CTA's smem, completion bytes signalled on an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2_both(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

}  // namespace ptx

template <int BN, int EPI>
struct Gemm2Smem {
  static constexpr int kABytes = BM * BK * 2;          // 128 rows of A per CTA: 16 KB
  static constexpr int kBBytes = (BN / 2) * BK * 2;    // half of the B tile per CTA
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytes = kEpiPairs * StagedEpi<EPI>::kBytesPerWarp;   // staging + aux boxes, one set per warp PAIR
  static constexpr int kBarrierBytes = 512;
  static constexpr int kAvail = 227 * 1024 - 1024 - kBarrierBytes - kEpiBytes;
#ifndef MTVAF_GEMM_MAX_STAGES
#define MTVAF_GEMM_MAX_STAGES 8            // experiments: -DMTVAF_GEMM_MAX_STAGES=n caps the operand ring
#endif
  static constexpr int kStages =
      (kAvail / kStageBytes) > MTVAF_GEMM_MAX_STAGES ? MTVAF_GEMM_MAX_STAGES : (kAvail / kStageBytes);
  static constexpr int kTotal = kStages * kStageBytes + kEpiBytes + kBarrierBytes + 1024;
  static_assert(kStages >= 3, "not enough shared memory for the operand ring");
};

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads2, 1)
gemm_bf16_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                     const __grid_constant__ CUtensorMap tmAux, const EpiArgs ep_in, int M, int N, int K, int splits,
                     int kb_per) {
  const EpiArgs ep = resolve_step(ep_in);
  using namespace ptx;
  using S = Gemm2Smem<BN, EPI>;
  using SE = StagedEpi<EPI>;
  constexpr int BM2 = 2 * BM;                          // rows per cluster tile
  constexpr int BNH = BN / 2;                          // B rows staged per CTA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* epi_smem = smem + S::kStages * S::kStageBytes;                 // per-warp staging / aux boxes
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + S::kEpiBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full = empty_bar + S::kStages;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint64_t* aux_bar = tmem_empty + 2;              // [kEpiPairs][2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aux_bar + 2 * kEpiPairs);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int n_tiles_m = (M + BM2 - 1) / BM2;
  const int n_tiles_n = (N + BN - 1) / BN;
  const int kb_total = (K + BK - 1) / BK;
  const int n_items = n_tiles_m * n_tiles_n * splits;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * kEpiWarps2);    // both CTAs' epilogue warps (used in the leader only)
    }
    for (int s = 0; s < 2 * kEpiPairs; ++s) mbar_init(&aux_bar[s], 1);
    fence_barrier_init();
  }
  constexpr int kTmemCols = 2 * BN;
  if (warp == 1) tmem_alloc_cg2<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  cluster_arrive();                                  // barriers of both CTAs initialised before any remote use
  cluster_wait();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const WorkItem w = decode_item(item, n_tiles_n, n_tiles_m, kb_total, kb_per);   // w.m0 in units of BM
        const int m0 = 2 * w.m0 + rank * BM;
        const int n0 = w.tn * BN + rank * BNH;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = tiles + stage * S::kStageBytes;
          uint8_t* sB = sA + S::kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
          const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
          const int k0 = kb * BK;
          if (!A_MN) {
            tma_load_2d_cg2(sA, &tmA, bar, k0, m0);                        // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)                              // boxes {64 m, 64 k}
              tma_load_2d_cg2(sA + j * 8192, &tmA, bar, m0 + j * 64, k0);
          }
          if (!B_MN) {
            tma_load_2d_cg2(sB, &tmB, bar, k0, n0);                        // box {64 k, BN/2 n}
          } else {
#pragma unroll
            for (int j = 0; j < BNH / 64; ++j)                             // boxes {64 n, 64 k}
              tma_load_2d_cg2(sB + j * 8192, &tmB, bar, n0 + j * 64, k0);
          }
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM2, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
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
            umma_f16_ss_cg2(d_tmem, da, db, idesc, (kb > w.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_cg2_both(&empty_bar[stage]);   // frees this stage in both CTAs when the MMAs retire
          if (++stage == S::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_cg2_both(&tmem_full[acc]);       // accumulator complete -> both CTAs' epilogues
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..17), both CTAs =====================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int sub = (warp - 2) >> 2;                 // 0..3: which quarter of the tile's columns
    const int half = sub >> 1;                       // which half of the tile's columns the warp PAIR owns
    const int part = sub & 1;                        // which 32 columns of each 64-column box this warp computes
    const int pair = half * 4 + quad;                // 0..7: owner of one set of staging / aux boxes
    const int bar_id = 1 + pair;                     // named barrier of the pair (64 threads)
    int acc = 0;
    uint32_t acc_phase = 0;
    if (SE::kStaged && ep.staged) {
      // ---- coalesced path: TMA-loaded aux boxes, smem-staged TMA stores (epilogue_staged.cuh)
      constexpr int kBoxes = BN / 128;               // 64-column boxes per warp pair per tile
      const bool issuer = part == 0 && lane == 0;    // the pair's TMA thread
      uint8_t* my = epi_smem + pair * SE::kBytesPerWarp;
      uint8_t* out_box = my;                         // [kOutBufs] boxes
      uint8_t* aux_box = my + SE::kOutBufs * kEpiBoxBytes;   // [kAuxBufs] boxes
      constexpr int kAB = SE::kAuxBufs > 0 ? SE::kAuxBufs : 1;
      uint64_t* my_bar = aux_bar + 2 * pair;
      // prefetch iterator over this pair's valid boxes (runs two boxes ahead of the consumer; issuer thread only)
      int pf_item = cluster_id, pf_j = -1;
      auto pf_next = [&](int& row0, int& col0) -> bool {
        while (true) {
          if (++pf_j == kBoxes) { pf_j = 0; pf_item += n_clusters; }
          if (pf_item >= n_items) return false;
          const WorkItem w = decode_item(pf_item, n_tiles_n, n_tiles_m, kb_total, kb_per);
          col0 = w.tn * BN + half * (BN / 2) + pf_j * 64;
          row0 = 2 * w.m0 + rank * BM + quad * 32;
          if (col0 < N) return true;
        }
      };
      uint32_t used = 0;                             // boxes consumed so far (aux buffer = used & 1)
      if (SE::kAux && issuer) {
        for (int i = 0; i < kAB; ++i) {
          int r0, c0;
          if (!pf_next(r0, c0)) break;
          mbar_arrive_expect_tx(&my_bar[i], kEpiBoxBytes);
          tma_load_2d(aux_box + i * kEpiBoxBytes, &tmAux, &my_bar[i], c0, r0);
        }
      }
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const WorkItem w = decode_item(item, n_tiles_n, n_tiles_m, kb_total, kb_per);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const int row0 = 2 * w.m0 + rank * BM + quad * 32;
        const int row = row0 + lane;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * (BN / 2);
#pragma unroll 1
        for (int j = 0; j < kBoxes; ++j) {
          const int col0 = w.tn * BN + half * (BN / 2) + j * 64;
          if (col0 >= N) continue;                   // pair-uniform
          const int c32 = col0 + part * 32;          // this warp's 32 columns
          const uint32_t ab = kAB == 2 ? (used & 1) : 0;
          const uint32_t ab_phase = kAB == 2 ? ((used >> 1) & 1) : (used & 1);
          const uint32_t aux_a = smem_u32(aux_box) + ab * kEpiBoxBytes;
          const uint32_t out_a = SE::kInplace ? aux_a : smem_u32(out_box);
          // two 16-column steps; the TMEM load and the bias of step 1 are in flight while step 0 is computed
          uint32_t ra[16], rb[16];
          tmem_ld_32x32b_x16(t_addr + j * 64 + part * 32, ra);
          float bias_a[16], bias_b[16];
          epi_load_bias16(ep, c32, N, c32 + 16 <= N, bias_a);
          uint32_t keep = 0xFFFFFFFFu;
          if (EPI == MTVAF_EPI_RESID && ep.drop_threshold)
            keep = dropout_mask32(ep.seed, (unsigned long long)row * (unsigned long long)N + c32, ep.drop_threshold);
          if (SE::kAux) mbar_wait(&my_bar[ab], ab_phase);
          uint4 aux2[2];
          if (SE::kAux) {
#pragma unroll
            for (int g = 0; g < 2; ++g) aux2[g] = ld_shared_v4(aux_a + box_piece_off(lane, part * 4 + g));
          }
          tmem_ld_wait();
          tmem_ld_32x32b_x16(t_addr + j * 64 + part * 32 + 16, rb);
          epi_load_bias16(ep, c32 + 16, N, c32 + 32 <= N, bias_b);
          uint32_t o[8], p[8];
          epi_compute_groups<EPI, 2>(ep, ra, bias_a, aux2, keep, o, p);
          if (!SE::kInplace) {
            // the pair's staging box is free once the previous TMA store has finished READING it
            if (issuer) tma_store_wait_read<0>();
            __syncwarp();
            named_bar_sync(bar_id, 64);
          }                                          // (in place: each thread overwrites exactly the aux pieces it read)
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t off = box_piece_off(lane, part * 4 + g);
            st_shared_v4(out_a + off, o[g * 4], o[g * 4 + 1], o[g * 4 + 2], o[g * 4 + 3]);
            if (SE::kTwoOut)
              st_shared_v4(out_a + kEpiBoxBytes + off, p[g * 4], p[g * 4 + 1], p[g * 4 + 2], p[g * 4 + 3]);
          }
          if (SE::kAux) {
#pragma unroll
            for (int g = 0; g < 2; ++g) aux2[g] = ld_shared_v4(aux_a + box_piece_off(lane, part * 4 + 2 + g));
          }
          tmem_ld_wait();
          epi_compute_groups<EPI, 2>(ep, rb, bias_b, aux2, keep >> 16, o, p);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t off = box_piece_off(lane, part * 4 + 2 + g);
            st_shared_v4(out_a + off, o[g * 4], o[g * 4 + 1], o[g * 4 + 2], o[g * 4 + 3]);
            if (SE::kTwoOut)
              st_shared_v4(out_a + kEpiBoxBytes + off, p[g * 4], p[g * 4 + 1], p[g * 4 + 2], p[g * 4 + 3]);
          }
          fence_proxy_async_smem();                  // staging writes (and aux reads) ordered before the TMA ops
          __syncwarp();
          named_bar_sync(bar_id, 64);                // both halves of the box staged, both warps done with the aux box
          if (issuer) {
            tma_store_2d(&tmOut, SE::kInplace ? aux_box + ab * kEpiBoxBytes : out_box, col0, row0);
            if (SE::kTwoOut && ep.out2) tma_store_2d(&tmOut2, out_box + kEpiBoxBytes, col0, row0);
            tma_store_commit();
            if (SE::kAux && !SE::kInplace) {         // refill the aux box just consumed with the next box
              int r0n, c0n;
              if (pf_next(r0n, c0n)) {
                mbar_arrive_expect_tx(&my_bar[ab], kEpiBoxBytes);
                tma_load_2d(aux_box + ab * kEpiBoxBytes, &tmAux, &my_bar[ab], c0n, r0n);
              }
            }
          }
          if (ep.colsum) {                           // kernel-uniform
            // column sums of the box just staged (the bf16 values as stored).  This warp: column pairs
            // [16 part, 16 part + 16); lanes 0-15 sum rows 0-15, lanes 16-31 rows 16-31, combined by one shuffle.
            // Rows past M hold epi(0) of zero-filled operand rows: excluded.  (The next box's staging writes come
            // after its first named barrier, i.e. after both warps' reads here.)
            const int rows_ok = min(32, M - row0);
            const int cp = (lane & 15) + 16 * part;  // column pair 0..31 of the box
            const int rbase = (lane >> 4) * 16;
            const uint32_t colb = out_a + ((cp & 3) << 2);
            const int piece = cp >> 2;
            uint32_t v[16];
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
              const int r_ = rbase + rr;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v[rr]) : "r"(colb + r_ * 128 + ((piece ^ (r_ & 7)) << 4)));
            }
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
              const float2 f = unpack_bf16x2(rbase + rr < rows_ok ? v[rr] : 0u);
              s0 += f.x; s1 += f.y;
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            const int cc = col0 + 2 * cp;
            if (lane < 16) {
              if (cc < N) atomicAdd(ep.colsum + cc, s0);
              if (cc + 1 < N) atomicAdd(ep.colsum + cc + 1, s1);
            }
          }
          if (SE::kInplace) {
            // the box is both the store's source and the next aux load's destination: the load is issued once the store
            // has finished reading it (and, with colsum, once both warps have read their column sums out of it)
            if (ep.colsum) {
              __syncwarp();
              named_bar_sync(bar_id, 64);
            }
            if (issuer) {
              int r0n, c0n;
              if (pf_next(r0n, c0n)) {
                tma_store_wait_read<0>();
                mbar_arrive_expect_tx(&my_bar[ab], kEpiBoxBytes);
                tma_load_2d(aux_box + ab * kEpiBoxBytes, &tmAux, &my_bar[ab], c0n, r0n);
              }
            }
            __syncwarp();
          }
          ++used;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (issuer) tma_store_wait<0>();               // all output bytes written before the CTA retires
      __syncwarp();
    } else {
      // ---- direct path (fp32 outputs, atomics, row reductions, unaligned outputs): a quarter of the columns per warp
      constexpr int kChunks = BN / 128;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const WorkItem w = decode_item(item, n_tiles_n, n_tiles_m, kb_total, kb_per);
        const int n0 = w.tn * BN + sub * (BN / 4);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const int row = 2 * w.m0 + rank * BM + quad * 32 + lane;
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + sub * (BN / 4);
        float rowacc = 0.f;
#pragma unroll 1
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
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  // teardown: nobody may leave (or free TMEM) while the pair still reads this CTA's smem / writes its TMEM
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_arrive();
  cluster_wait();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2<kTmemCols>(tmem_base);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch_gemm_tc2(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, const EpiArgs& ep,
                    int splits, cudaStream_t stream) {
  using S = Gemm2Smem<BN, EPI>;
  using SE = StagedEpi<EPI>;
  CUtensorMap tmA, tmB, tmOut, tmOut2, tmAux;
  int rc;
  if (!A_MN) rc = make_tmap_bf16_2d(&tmA, A, K, M, lda, BK, BM);
  else       rc = make_tmap_bf16_2d(&tmA, A, M, K, lda, 64, BK);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_bf16_2d(&tmB, B, K, N, ldb, BK, BN / 2);
  else       rc = make_tmap_bf16_2d(&tmB, B, N, K, ldb, 64, BK);
  if (rc) return rc;

  EpiArgs epl = ep;
  tmOut = tmA; tmOut2 = tmA; tmAux = tmA;            // placeholders when unused
  if (SE::kStaged && epl.staged) {
    // [32 rows x 64 cols] bf16 boxes of the output (and pre-activation / aux operands)
    rc = make_tmap_bf16_2d(&tmOut, epl.out, N, M, epl.ldo, 64, 32);
    if (rc) return rc;
    if (SE::kTwoOut && epl.out2) {
      rc = make_tmap_bf16_2d(&tmOut2, epl.out2, N, M, epl.ld_out2, 64, 32);
      if (rc) return rc;
    }
    if (SE::kAux) {
      rc = make_tmap_bf16_2d(&tmAux, epl.aux, N, M, epl.ld_aux, 64, 32);
      if (rc) return rc;
    }
  } else {
    epl.staged = 0;
  }
  const int kb_total = (K + BK - 1) / BK;
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  const int kb_per = (kb_total + splits - 1) / splits;
  splits = (kb_total + kb_per - 1) / kb_per;
  const int n_items = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN) * splits;
  const int max_clusters = sm_count() / 2;
  const int clusters = n_items < max_clusters ? n_items : max_clusters;

  auto kern = gemm_bf16_tc2_kernel<BN, A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  kern<<<2 * clusters, kGemmThreads2, S::kTotal, stream>>>(tmA, tmB, tmOut, tmOut2, tmAux, epl, M, N, K, splits, kb_per);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

// 0 = auto (CTA pairs when the problem has at least one full pair tile), 1 = single-CTA kernel only
int gemm_impl_override();

// dispatch: CTA-pair kernel for everything with M >= 256 (the training GEMMs), single-CTA kernel otherwise
#define MTVAF_GEMM_CASE2(MODE_, AMN_, BMN_)                                                                \
  case MODE_:                                                                                              \
    if (pair) return narrow ? launch_gemm_tc2<128, AMN_, BMN_, MODE_>(A, lda, B, ldb, M, N, K, ep, splits, stream) \
                            : launch_gemm_tc2<256, AMN_, BMN_, MODE_>(A, lda, B, ldb, M, N, K, ep, splits, stream); \
    return narrow ? launch_gemm_tc<128, AMN_, BMN_, MODE_>(A, lda, B, ldb, M, N, K, ep, splits, stream)    \
                  : launch_gemm_tc<256, AMN_, BMN_, MODE_>(A, lda, B, ldb, M, N, K, ep, splits, stream)

}  // namespace mtvaf
