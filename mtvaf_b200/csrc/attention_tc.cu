// Prefix ("fusion") self-attention on the tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
//
// PERSISTENT CTAs (256 threads, two per SM so one CTA's softmax overlaps the other's MMAs / loads) walk the
// (batch, head, 128-query tile) items.  Per item:
//   TMA  : Q tile, the visual-prefix K_p/V_p rows (8-row boxes) and the text K/V rows (64-row boxes)
//          land in SWIZZLE_128B shared memory -- the torch.cat of models/modeling_roberta.py:221-222 is
//          just two groups of TMA boxes into one key-row numbering (prefix padded to a multiple of 8);
//          the next item's Q / K loads are issued as soon as S has retired, its V load as soon as O has retired
//   MMA 1: S[q, key] = Q K^T  (tcgen05.mma, both operands K-major)            -> TMEM columns [0, N16)
//   SIMT : thread = (query row == TMEM lane, every other 8-key unit): 1/sqrt(d) and the additive key mask
//          (-10000.0, models/modeling_roberta.py:1000) folded into one FFMA in the log2 domain, row max and
//          row sum exchanged between the two threads of a row through shared memory, ex2.approx, optional
//          dropout (one hash per PAIR of keys), P written as bf16 into shared memory in the K-major
//          SWIZZLE_128B operand layout (one 16-byte piece per unit)
//   MMA 2: O[q, d] = P V      (A = P K-major from smem, B = V MN-major)       -> TMEM columns [S_COLS, +64)
//   store: O / rowsum -> ctx (heads merged, :276-278; 64 B per thread), lse = max + log(rowsum) for backward.
// The backward kernel (same tiling, five MMAs) is in attention_tc_bwd.cu.
#include "attention_tc.cuh"
#include <cstdlib>
#include <algorithm>

namespace mtvaf {
using namespace ptx;

constexpr int kFwdThreadsSmall = 256, kFwdThreadsBig = 512;   // 2 / 4 threads per query row
constexpr float kFwdLog2e = 1.4426950408889634f;

// Single-thread work (TMA and tcgen05.mma issue) sits behind `warp == 0 && elect_one()`, not `tid == 0`: from a branch on the
// thread index the compiler wraps EVERY tcgen05.mma / TMA instruction in an ELECT / R2UR / BRA.U.ANY loop over the active
// lanes (~10 instructions and ~55 cycles per MMA); behind elect.sync it issues them back to back (tools/micro/umma_rate.cu:
// 55 -> 44 cycles per M128 N64 MMA at issue, retire floor 59; in these kernels ~95 -> ~60).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// BIG (N16 > 192: one CTA per SM) runs 16 warps = four threads per query row (+9 % at L = 512, +8 % at P = 100, measured
// round 2); the small shapes (two CTAs per SM) are faster with 8 warps = two threads per row.
template <bool BIG>
__global__ void __launch_bounds__(BIG ? kFwdThreadsBig : kFwdThreadsSmall, BIG ? 1 : 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmKp, const __grid_constant__ CUtensorMap tmVp,
                   AttnTcArgs a, __nv_bfloat16* __restrict__ ctx, long long ld_ctx, float* __restrict__ lse_out) {
  constexpr int kFwdThreads = BIG ? kFwdThreadsBig : kFwdThreadsSmall;
  constexpr int NPART = kFwdThreads / 128;           // threads per query row
  constexpr int S_COLS = BIG ? 448 : 192;            // TMEM columns reserved for S; O follows
  constexpr int TMEM_COLS = BIG ? 512 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int key_rows = a.P8 + a.L64;                 // smem key rows (multiples of 8 / 64)
  const int n_chunks = (a.N16 + 63) / 64;            // 64-key chunks of P
  uint8_t* sQ = smem;                                // [128][64] bf16, 16 KB
  uint8_t* sK = sQ + 16384;                          // [key_rows][64]
  uint8_t* sV = sK + key_rows * 128;
  uint8_t* sP = sV + key_rows * 128;                 // n_chunks x [128][64] bf16
  float* sMask = reinterpret_cast<float*>(sP + n_chunks * 16384);     // [N16], additive mask * log2(e)
  float* sExch = sMask + ((a.N16 + 15) / 16) * 16;   // [2][4][128]: partial row max / row sum per part of the units
  uint64_t* bars = reinterpret_cast<uint64_t*>(sExch + 1024);         // Q+K landed, s, o, V landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;       // TMEM lane group / which part of the units (0..NPART-1)
  const int row = quad * 32 + lane;
  const int H = a.nh * 64;
  const int q_tiles = (a.L + 127) / 128;
  const int n_items = a.B * a.nh * q_tiles;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmKV);
    if (a.P8 > 0) { prefetch_tmap(&tmKp); prefetch_tmap(&tmVp); }
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
    fence_barrier_init();
  }
  __syncwarp();                                      // tcgen05.alloc is .sync.aligned: reconverge warp 0 first
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Q and K are free again as soon as S = Q K^T has retired, V only after O = P V: two barriers, so the next item's
  // Q / K tiles travel during this item's softmax and its S MMA can start the moment the item begins
  auto issue_qk = [&](int item) {                    // one thread
    const int qt = item % q_tiles, bh = item / q_tiles;
    const int b = bh / a.nh, h = bh - b * a.nh;
    mbar_arrive_expect_tx(&bars[0], 16384u + (uint32_t)key_rows * 128u);
    tma_load_2d(sQ, &tmQ, &bars[0], h * 64, b * a.L + qt * 128);
    for (int r = 0; r < a.P8; r += 8) tma_load_2d(sK + r * 128, &tmKp, &bars[0], 0, (b * a.nh + h) * a.P + r);
    for (int r = 0; r < a.L64; r += 64)
      tma_load_2d(sK + (a.P8 + r) * 128, &tmKV, &bars[0], H + h * 64, b * a.L + a.kt0 + r);
  };
  auto issue_v = [&](int item) {                     // one thread
    const int bh = item / q_tiles;
    const int b = bh / a.nh, h = bh - b * a.nh;
    mbar_arrive_expect_tx(&bars[3], (uint32_t)key_rows * 128u);
    for (int r = 0; r < a.P8; r += 8) tma_load_2d(sV + r * 128, &tmVp, &bars[3], 0, (b * a.nh + h) * a.P + r);
    for (int r = 0; r < a.L64; r += 64)
      tma_load_2d(sV + (a.P8 + r) * 128, &tmKV, &bars[3], 2 * H + h * 64, b * a.L + a.kt0 + r);
  };
  if (warp == 0 && (int)blockIdx.x < n_items && elect_one()) { issue_qk(blockIdx.x); issue_v(blockIdx.x); }

  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  const float sc2 = a.scale * kFwdLog2e;             // fold log2(e): exp(x) = exp2(x * log2e)
  const int row8 = row & 7;
  const uint32_t prow_off = (row >> 3) * 1024 + row8 * 128;
  const int units = a.N16 >> 3;
  const uint32_t aMask = smem_u32(sMask), aPs = smem_u32(sP), aExch = smem_u32(sExch);   // explicit shared-space accesses

  // key `k` -> additive mask (x log2 e); keys >= N16 are never read (N16 <= 448 < 2 * kFwdThreads).  Only text keys inside
  // the window depend on memory: the raw key-mask word is PREFETCHED one item ahead and nothing is computed from it until
  // the next item stores the mask, so the load has a whole item to land (computing the value right after the load stalled
  // every warp on the long scoreboard once per item)
  auto key_const = [&](int k) -> float {
    if (k >= a.N16) return 0.f;
    if (k < a.P8) return (k < a.P) ? 0.f : -INFINITY;
    return -INFINITY;                                  // text key outside the window (inside: from the key mask)
  };
  const bool dyn0 = tid >= a.P8 && tid < a.N16 && tid - a.P8 < a.Lk;
  const bool dyn1 = tid + kFwdThreads >= a.P8 && tid + kFwdThreads < a.N16 && tid + kFwdThreads - a.P8 < a.Lk;
  const float const0 = key_const(tid), const1 = key_const(tid + kFwdThreads);
  auto fetch_raw = [&](int b, bool dyn, int k) -> long long {
    return dyn ? __ldg(a.key_mask + (long long)b * a.L + a.kt0 + (k - a.P8)) : 1;
  };
  long long km_next[2] = {1, 1};
  if ((int)blockIdx.x < n_items) {
    const int b0 = (int)blockIdx.x / q_tiles / a.nh;
    km_next[0] = fetch_raw(b0, dyn0, tid);
    km_next[1] = fetch_raw(b0, dyn1, tid + kFwdThreads);
  }
  const unsigned long long seed_eff = a.drop_thr ? step_seed(a.seed, a.step) : 0ull;

  uint32_t ph = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ph ^= 1) {
    const int qt = item % q_tiles, bh = item / q_tiles;
    const int b = bh / a.nh, h = bh - b * a.nh;
    const int q = qt * 128 + row;
    // additive key mask (x log2 e) in smem-key numbering: prefix rows [0,P) visible, [P,P8) padding, text rows follow
    if (tid < a.N16) sMask[tid] = dyn0 ? (km_next[0] != 0 ? 0.f : -10000.0f * kFwdLog2e) : const0;
    if (tid + kFwdThreads < a.N16) sMask[tid + kFwdThreads] = dyn1 ? (km_next[1] != 0 ? 0.f : -10000.0f * kFwdLog2e) : const1;
    if (item + (int)gridDim.x < n_items) {
      const int bn = (item + (int)gridDim.x) / q_tiles / a.nh;
      km_next[0] = fetch_raw(bn, dyn0, tid);
      km_next[1] = fetch_raw(bn, dyn1, tid + kFwdThreads);
    }
    __syncwarp();                                      // elect.sync needs the whole warp
    if (warp == 0 && elect_one()) {
      mbar_wait(&bars[0], ph);
      tc_fence_after();
      // ---- S = Q K^T
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
      for (int n0 = 0; n0 < a.N16; n0 += 256) {
        const int n = min(256, a.N16 - n0);
        const uint32_t idesc = make_idesc_bf16(128, n, false, false);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_base + n0, make_smem_desc_sw128(aQ + k * 32, 16, 1024),
                      make_smem_desc_sw128(aK + n0 * 128 + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
      }
      umma_commit(&bars[1]);
    }
    __syncthreads();                                   // publishes sMask (the MMA is in flight meanwhile)
    mbar_wait(&bars[1], ph);
    __syncwarp();
    tc_fence_after();
    if (warp == 0 && item + (int)gridDim.x < n_items && elect_one()) issue_qk(item + gridDim.x);   // S retired: Q / K are free
    __syncwarp();

    // ---- pass 1: row max of (S / sqrt(d) + mask) * log2 e over this thread's units
    float mx = -INFINITY;
    for (int u = half; u < units; u += NPART) {
      const int c = u << 3;
      uint32_t r[8];
      tmem_ld_32x32b_x8(t_row + c, r);
      const float4 m0 = lds_f4(aMask + c * 4);
      const float4 m1 = lds_f4(aMask + c * 4 + 16);
      tmem_ld_wait();
      mx = fmaxf(mx, fmaxf(fmaxf(fmaf(__uint_as_float(r[0]), sc2, m0.x), fmaf(__uint_as_float(r[1]), sc2, m0.y)),
                           fmaxf(fmaf(__uint_as_float(r[2]), sc2, m0.z), fmaf(__uint_as_float(r[3]), sc2, m0.w))));
      mx = fmaxf(mx, fmaxf(fmaxf(fmaf(__uint_as_float(r[4]), sc2, m1.x), fmaf(__uint_as_float(r[5]), sc2, m1.y)),
                           fmaxf(fmaf(__uint_as_float(r[6]), sc2, m1.z), fmaf(__uint_as_float(r[7]), sc2, m1.w))));
    }
    sts_f32(aExch + (half * 128 + row) * 4, mx);
    __syncthreads();
    // finite: key 0 exists and S is finite (a part without units contributes -inf)
    float mx2 = fmaxf(lds_f32(aExch + row * 4), lds_f32(aExch + (128 + row) * 4));
    if (NPART == 4) mx2 = fmaxf(mx2, fmaxf(lds_f32(aExch + (256 + row) * 4), lds_f32(aExch + (384 + row) * 4)));

    // ---- pass 2: P = exp2(x - max), row sum, dropout, bf16 P -> shared memory.  With dropout the 1/(1-p) scale
    // rides in the exponent (max - log2(scale)): P is born scaled, the row sum is corrected once at the end.
    const uint32_t rowkey =
        a.drop_thr ? attn_drop_rowkey(seed_eff, ((unsigned long long)b * a.nh + h) * a.L + q) : 0u;
    const float mxs = a.drop_thr ? mx2 - log2f(a.drop_scale) : mx2;
    float sum = 0.f;
    for (int u = half; u < units; u += NPART) {
      const int c = u << 3;
      uint32_t r[8];
      tmem_ld_32x32b_x8(t_row + c, r);
      const float4 m0 = lds_f4(aMask + c * 4);
      const float4 m1 = lds_f4(aMask + c * 4 + 16);
      const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      tmem_ld_wait();
      float p[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        p[j] = ex2_approx(fmaf(__uint_as_float(r[j]), sc2, mk[j] - mxs));
        sum += p[j];
      }
      if (a.drop_thr) attn_drop_apply8(rowkey, c < a.P8 ? c : a.kbase + (c - a.P8), a.drop_thr, p);
      uint4 w;
      w.x = pack_bf16x2(p[0], p[1]); w.y = pack_bf16x2(p[2], p[3]);
      w.z = pack_bf16x2(p[4], p[5]); w.w = pack_bf16x2(p[6], p[7]);
      sts_u4(aPs + (u >> 3) * 16384 + prow_off + (((u & 7) ^ row8) << 4), w);
    }
    sts_f32(aExch + (512 + half * 128 + row) * 4, sum);
    fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0 && elect_one()) {
      // ---- O = P V
      mbar_wait(&bars[3], ph);                         // V landed (long ago: requested when the previous O retired)
      tc_fence_after();
      const uint32_t aP = smem_u32(sP), aV = smem_u32(sV);
      const uint32_t idesc = make_idesc_bf16(128, 64, false, true);
      const int ksteps = a.N16 / 16;
      for (int j = 0; j < ksteps; ++j)
        umma_f16_ss(tmem_base + S_COLS, make_smem_desc_sw128(aP + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(aV + j * 2048, 8192, 1024), idesc, j > 0 ? 1u : 0u);
      umma_commit(&bars[2]);
    }
    __syncwarp();
    float total = lds_f32(aExch + (512 + row) * 4) + lds_f32(aExch + (640 + row) * 4);
    if (NPART == 4) total += lds_f32(aExch + (768 + row) * 4) + lds_f32(aExch + (896 + row) * 4);
    mbar_wait(&bars[2], ph);
    __syncwarp();
    tc_fence_after();
    // V / P of this item are consumed: fetch the next item's V behind the epilogue
    if (warp == 0 && item + (int)gridDim.x < n_items && elect_one()) issue_v(item + gridDim.x);
    __syncwarp();

    // tcgen05.ld is warp-collective (.sync.aligned): EVERY lane issues the loads (rows past L included,
    // never inside a divergent branch); only the global stores are predicated.
    // O = (sum_k keep_k P'_k V_k) / (sum_k P'_k / scale) with P' = scale * P
    const float inv = a.drop_scale / total;
    const bool valid = q < a.L;
    if constexpr (NPART == 4) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_row + S_COLS + half * 16, r);
      tmem_ld_wait();
      if (valid) {
        uint4* o = reinterpret_cast<uint4*>(ctx + ((long long)b * a.L + q) * ld_ctx + h * 64 + half * 16);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]) * inv, __uint_as_float(r[v * 8 + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]) * inv, __uint_as_float(r[v * 8 + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]) * inv, __uint_as_float(r[v * 8 + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]) * inv, __uint_as_float(r[v * 8 + 7]) * inv);
          o[v] = w;
        }
      }
    } else {
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + S_COLS + half * 32, r);
      tmem_ld_wait();
      if (valid) {
        uint4* o = reinterpret_cast<uint4*>(ctx + ((long long)b * a.L + q) * ld_ctx + h * 64 + half * 32);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]) * inv, __uint_as_float(r[v * 8 + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]) * inv, __uint_as_float(r[v * 8 + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]) * inv, __uint_as_float(r[v * 8 + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]) * inv, __uint_as_float(r[v * 8 + 7]) * inv);
          o[v] = w;
        }
      }
    }
    if (valid && half == 0)
      lse_out[((long long)b * a.nh + h) * a.L + q] = (mxs + log2f(total)) * 0.6931471805599453f;
    // TMEM reads done before the next item's MMAs overwrite S / O; orders sMask / sExch reuse
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---- short text (L <= 64): TWO (batch, head) items per 128-row tile -------------------------------------------------
// A 128-row tile holds only L <= 64 query rows of one item: half the TMEM lanes (= half the softmax threads) would work
// on rows nobody stores.  Here rows [0, 64) belong to item 2i and rows [64, 128) to item 2i + 1; the keys of the two
// items are laid end to end (KR = P8 + 64 smem rows each), S = [Q_a; Q_b] [K_a; K_b]^T is ONE M = 128, N = 2 KR MMA whose
// off-diagonal blocks are simply never read, and P is block diagonal -- its off-diagonal pieces in shared memory are
// zeroed once per CTA and never written again -- so O = P [V_a; V_b] is again one MMA.  The tensor pipe does twice the
// useful flops (it idles at < 10 % either way); every softmax thread now works on a row that exists.
__global__ void __launch_bounds__(kFwdThreadsSmall, 2)
attn_fwd_tc_pair_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmKp,
                        const __grid_constant__ CUtensorMap tmVp, AttnTcArgs a, __nv_bfloat16* __restrict__ ctx,
                        long long ld_ctx, float* __restrict__ lse_out) {
  constexpr int S_COLS = 192, TMEM_COLS = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KR = a.P8 + 64;                          // key rows (= S columns) per item
  const int n_keys = 2 * KR;                         // multiple of 16
  const int n_chunks = (n_keys + 63) / 64;
  uint8_t* sQ = smem;                                // [128][64] bf16: rows [0,64) item a, [64,128) item b
  uint8_t* sK = sQ + 16384;                          // [2 KR][64]
  uint8_t* sV = sK + n_keys * 128;
  uint8_t* sP = sV + n_keys * 128;                   // n_chunks x [128][64] bf16, block diagonal
  float* sMask = reinterpret_cast<float*>(sP + n_chunks * 16384);     // [2 KR]
  float* sExch = sMask + ((n_keys + 15) / 16) * 16;  // [2][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sExch + 512);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const int blk = quad >> 1;                         // which item of the pair this row belongs to (warp-uniform)
  const int H = a.nh * 64;
  const int n_bh = a.B * a.nh;
  const int n_items = (n_bh + 1) / 2;

  if (tid == 0) {
    prefetch_tmap(&tmKV);
    if (a.P8 > 0) { prefetch_tmap(&tmKp); prefetch_tmap(&tmVp); }
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_ptr);
  // P starts as zeros: the off-diagonal pieces stay that way for the life of the CTA
  for (int i = tid; i < n_chunks * 1024; i += kFwdThreadsSmall) sts_u4(smem_u32(sP) + i * 16, make_uint4(0, 0, 0, 0));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // an odd item count leaves the last pair with one item: its second half re-reads the first (never stored)
  auto bh_of = [&](int item, int j) { const int bh = 2 * item + j; return bh < n_bh ? bh : 2 * item; };
  auto issue_qk = [&](int item) {                    // one thread
    mbar_arrive_expect_tx(&bars[0], 16384u + (uint32_t)n_keys * 128u);
    for (int j = 0; j < 2; ++j) {
      const int bh = bh_of(item, j), b = bh / a.nh, h = bh - b * a.nh;
      tma_load_2d(sQ + j * 8192, &tmKV, &bars[0], h * 64, b * a.L);
      uint8_t* k = sK + j * KR * 128;
      for (int r = 0; r < a.P8; r += 8) tma_load_2d(k + r * 128, &tmKp, &bars[0], 0, bh * a.P + r);
      tma_load_2d(k + a.P8 * 128, &tmKV, &bars[0], H + h * 64, b * a.L);
    }
  };
  auto issue_v = [&](int item) {                     // one thread
    mbar_arrive_expect_tx(&bars[3], (uint32_t)n_keys * 128u);
    for (int j = 0; j < 2; ++j) {
      const int bh = bh_of(item, j), b = bh / a.nh, h = bh - b * a.nh;
      uint8_t* v = sV + j * KR * 128;
      for (int r = 0; r < a.P8; r += 8) tma_load_2d(v + r * 128, &tmVp, &bars[3], 0, bh * a.P + r);
      tma_load_2d(v + a.P8 * 128, &tmKV, &bars[3], 2 * H + h * 64, b * a.L);
    }
  };
  if (warp == 0 && (int)blockIdx.x < n_items && elect_one()) { issue_qk(blockIdx.x); issue_v(blockIdx.x); }

  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  const float sc2 = a.scale * kFwdLog2e;
  const int row8 = row & 7;
  const uint32_t prow_off = (row >> 3) * 1024 + row8 * 128;
  const int units = KR >> 3;                         // per item
  const int u0 = blk * units;                        // first unit (8 S columns) of this row's item
  const uint32_t aMask = smem_u32(sMask), aPs = smem_u32(sP), aExch = smem_u32(sExch);

  // S column c = tid of the pair -> additive mask (x log2 e); 2 KR <= 192 < kFwdThreadsSmall.  The raw key-mask word is
  // prefetched one item ahead and only turned into a value when the next item stores it (see the kernel above)
  const int mj = tid >= KR ? 1 : 0, mk_ = tid - mj * KR;               // item of the pair / key within it
  const bool m_dyn = tid < n_keys && mk_ >= a.P8 && mk_ - a.P8 < a.L;
  const float m_const = tid >= n_keys ? 0.f : (mk_ < a.P8 ? (mk_ < a.P ? 0.f : -INFINITY) : -INFINITY);
  auto fetch_raw = [&](int item) -> long long {
    return m_dyn ? __ldg(a.key_mask + (long long)(bh_of(item, mj) / a.nh) * a.L + (mk_ - a.P8)) : 1;
  };
  long long km_next = (int)blockIdx.x < n_items ? fetch_raw(blockIdx.x) : 1;
  const unsigned long long seed_eff = a.drop_thr ? step_seed(a.seed, a.step) : 0ull;

  uint32_t ph = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ph ^= 1) {
    const int bh = 2 * item + blk;
    const bool exists = bh < n_bh;
    const int b = bh / a.nh, h = bh - b * a.nh;
    const int q = row & 63;
    if (tid < n_keys) sMask[tid] = m_dyn ? (km_next != 0 ? 0.f : -10000.0f * kFwdLog2e) : m_const;
    if (item + (int)gridDim.x < n_items) km_next = fetch_raw(item + gridDim.x);
    __syncwarp();                                      // elect.sync needs the whole warp
    if (warp == 0 && elect_one()) {
      mbar_wait(&bars[0], ph);
      tc_fence_after();
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
      const uint32_t idesc = make_idesc_bf16(128, n_keys, false, false);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tmem_base, make_smem_desc_sw128(aQ + k * 32, 16, 1024), make_smem_desc_sw128(aK + k * 32, 16, 1024),
                    idesc, k > 0 ? 1u : 0u);
      umma_commit(&bars[1]);
    }
    __syncthreads();
    mbar_wait(&bars[1], ph);
    __syncwarp();
    tc_fence_after();
    if (warp == 0 && item + (int)gridDim.x < n_items && elect_one()) issue_qk(item + gridDim.x);
    __syncwarp();

    float mx = -INFINITY;
    for (int u = half; u < units; u += 2) {
      const int c = (u0 + u) << 3;
      uint32_t r[8];
      tmem_ld_32x32b_x8(t_row + c, r);
      const float4 m0 = lds_f4(aMask + c * 4);
      const float4 m1 = lds_f4(aMask + c * 4 + 16);
      tmem_ld_wait();
      mx = fmaxf(mx, fmaxf(fmaxf(fmaf(__uint_as_float(r[0]), sc2, m0.x), fmaf(__uint_as_float(r[1]), sc2, m0.y)),
                           fmaxf(fmaf(__uint_as_float(r[2]), sc2, m0.z), fmaf(__uint_as_float(r[3]), sc2, m0.w))));
      mx = fmaxf(mx, fmaxf(fmaxf(fmaf(__uint_as_float(r[4]), sc2, m1.x), fmaf(__uint_as_float(r[5]), sc2, m1.y)),
                           fmaxf(fmaf(__uint_as_float(r[6]), sc2, m1.z), fmaf(__uint_as_float(r[7]), sc2, m1.w))));
    }
    sts_f32(aExch + (half * 128 + row) * 4, mx);
    __syncthreads();
    const float mx2 = fmaxf(lds_f32(aExch + row * 4), lds_f32(aExch + (128 + row) * 4));

    const uint32_t rowkey =
        a.drop_thr ? attn_drop_rowkey(seed_eff, ((unsigned long long)b * a.nh + h) * a.L + q) : 0u;
    const float mxs = a.drop_thr ? mx2 - log2f(a.drop_scale) : mx2;
    float sum = 0.f;
    for (int u = half; u < units; u += 2) {
      const int cl = u << 3, c = (u0 + u) << 3;
      uint32_t r[8];
      tmem_ld_32x32b_x8(t_row + c, r);
      const float4 m0 = lds_f4(aMask + c * 4);
      const float4 m1 = lds_f4(aMask + c * 4 + 16);
      const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      tmem_ld_wait();
      float p[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        p[j] = ex2_approx(fmaf(__uint_as_float(r[j]), sc2, mk[j] - mxs));
        sum += p[j];
      }
      if (a.drop_thr) attn_drop_apply8(rowkey, cl < a.P8 ? cl : a.kbase + (cl - a.P8), a.drop_thr, p);
      uint4 w;
      w.x = pack_bf16x2(p[0], p[1]); w.y = pack_bf16x2(p[2], p[3]);
      w.z = pack_bf16x2(p[4], p[5]); w.w = pack_bf16x2(p[6], p[7]);
      const int ug = u0 + u;
      sts_u4(aPs + (ug >> 3) * 16384 + prow_off + (((ug & 7) ^ row8) << 4), w);
    }
    sts_f32(aExch + (256 + half * 128 + row) * 4, sum);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0 && elect_one()) {
      mbar_wait(&bars[3], ph);
      tc_fence_after();
      const uint32_t aP = smem_u32(sP), aV = smem_u32(sV);
      const uint32_t idesc = make_idesc_bf16(128, 64, false, true);
      const int ksteps = n_keys / 16;
      for (int j = 0; j < ksteps; ++j)
        umma_f16_ss(tmem_base + S_COLS, make_smem_desc_sw128(aP + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(aV + j * 2048, 8192, 1024), idesc, j > 0 ? 1u : 0u);
      umma_commit(&bars[2]);
    }
    __syncwarp();
    const float total = lds_f32(aExch + (256 + row) * 4) + lds_f32(aExch + (384 + row) * 4);
    mbar_wait(&bars[2], ph);
    __syncwarp();
    tc_fence_after();
    if (warp == 0 && item + (int)gridDim.x < n_items && elect_one()) issue_v(item + gridDim.x);
    __syncwarp();

    const float inv = a.drop_scale / total;
    const bool valid = exists && q < a.L;
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_row + S_COLS + half * 32, r);
    tmem_ld_wait();
    if (valid) {
      uint4* o = reinterpret_cast<uint4*>(ctx + ((long long)b * a.L + q) * ld_ctx + h * 64 + half * 32);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]) * inv, __uint_as_float(r[v * 8 + 1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]) * inv, __uint_as_float(r[v * 8 + 3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]) * inv, __uint_as_float(r[v * 8 + 5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]) * inv, __uint_as_float(r[v * 8 + 7]) * inv);
        o[v] = w;
      }
      if (half == 0) lse_out[((long long)b * a.nh + h) * a.L + q] = (mxs + log2f(total)) * 0.6931471805599453f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

static size_t attn_fwd_tc_pair_smem(const AttnTcArgs& a) {
  const int n_keys = 2 * (a.P8 + 64);
  const int n_chunks = (n_keys + 63) / 64;
  return 1024 + 16384 + 2 * (size_t)n_keys * 128 + (size_t)n_chunks * 16384 + ((n_keys + 15) / 16) * 64 + 2048 + 64;
}

// two items per tile pay when a tile would otherwise be half empty and two CTAs still fit an SM
static bool attn_fwd_tc_pair_ok(const AttnTcArgs& a) {
  static const char* env = getenv("MTVAF_ATTN_FWD_PAIR");                  // experiment knob: 0 = off
  if (env && atoi(env) == 0) return false;
  return a.L <= 64 && a.kt0 == 0 && a.Lk == a.L && 2 * (a.P8 + 64) <= 192 && a.B * a.nh >= 2 &&
         2 * attn_fwd_tc_pair_smem(a) <= 227 * 1024;
}

size_t attn_fwd_tc_smem(const AttnTcArgs& a) {
  const int key_rows = a.P8 + a.L64;
  const int n_chunks = (a.N16 + 63) / 64;
  return 1024 + 16384 + 2 * (size_t)key_rows * 128 + (size_t)n_chunks * 16384 + ((a.N16 + 15) / 16) * 64 + 4096 + 64;
}

int attn_fwd_tc_launch(const AttnTcArgs& a, const AttnTcMaps& m, void* ctx, int64_t ld_ctx, float* lse,
                       cudaStream_t st) {
  const size_t smem = attn_fwd_tc_smem(a);
  const int n_items = a.B * a.nh * ((a.L + 127) / 128);
  const bool big = a.N16 > 192;
  static bool set0 = false, set1 = false, set2 = false;
  if (attn_fwd_tc_pair_ok(a)) {
    if (!set2) {
      MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set2 = true;
    }
    const int n_pairs = (a.B * a.nh + 1) / 2;
    const int grid = n_pairs < 2 * sm_count() ? n_pairs : 2 * sm_count();
    attn_fwd_tc_pair_kernel<<<grid, kFwdThreadsSmall, attn_fwd_tc_pair_smem(a), st>>>(m.kv, m.kp, m.vp, a, (__nv_bfloat16*)ctx, ld_ctx, lse);
    MTVAF_LAUNCH_CHECK();
    return 0;
  }
  if (!big) {
    if (!set0) {
      MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set0 = true;
    }
    int per_sm = 2 * smem <= 227 * 1024 ? 2 : 1;
    static const char* env_ps = getenv("MTVAF_ATTN_FWD_CTAS_PER_SM");      // experiment knob
    if (env_ps && atoi(env_ps) == 1) per_sm = 1;
    const int grid = n_items < per_sm * sm_count() ? n_items : per_sm * sm_count();
    attn_fwd_tc_kernel<false><<<grid, kFwdThreadsSmall, smem, st>>>(m.q, m.kv, m.kp, m.vp, a, (__nv_bfloat16*)ctx, ld_ctx, lse);
  } else {
    if (!set1) {
      MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set1 = true;
    }
    const int grid = n_items < sm_count() ? n_items : sm_count();
    attn_fwd_tc_kernel<true><<<grid, kFwdThreadsBig, smem, st>>>(m.q, m.kv, m.kp, m.vp, a, (__nv_bfloat16*)ctx, ld_ctx, lse);
  }
  MTVAF_LAUNCH_CHECK();
  return 0;
}

bool attn_fwd_tc_fits(const AttnTcArgs& a) { return a.N16 <= 448 && attn_fwd_tc_smem(a) <= 227 * 1024; }

// ---- long text: two key windows + merge -----------------------------------------------------------------------------
// softmax over the union of two key sets from the two partial results: with lse_i = log sum_{k in window i} exp(s_k),
//   lse = log(exp(lse_0) + exp(lse_1)),   O = exp(lse_0 - lse) O_0 + exp(lse_1 - lse) O_1
// (dropout sits in the numerators O_i only, so the merge is unaffected).  One thread per (token, head, 8 columns).
__global__ void __launch_bounds__(256)
attn_merge_windows_kernel(__nv_bfloat16* __restrict__ ctx, long long ld_ctx, float* __restrict__ lse,
                          const __nv_bfloat16* __restrict__ ctx1, const float* __restrict__ lse1, int B, int L, int nh) {
  const long long n = (long long)B * L * nh * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i & 7);
    const long long th = i >> 3;
    const int h = (int)(th % nh);
    const long long t = th / nh;                       // token = b * L + q
    const long long b = t / L, q = t - b * L;
    const long long li = (b * nh + h) * L + q;
    const float l0 = lse[li], l1 = lse1[li];
    const float mx = fmaxf(l0, l1);
    const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx);
    const float inv = 1.f / (e0 + e1);
    const float w0 = e0 * inv, w1 = e1 * inv;
    __nv_bfloat16* o = ctx + t * ld_ctx + h * 64 + v * 8;
    const __nv_bfloat16* o1 = ctx1 + (t * nh + h) * 64 + v * 8;
    float a0[8], a1[8];
    Vec8<__nv_bfloat16>::load(o, a0);
    Vec8<__nv_bfloat16>::load(o1, a1);
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = fmaf(w0, a0[j], w1 * a1[j]);
    Vec8<__nv_bfloat16>::store(o, a0);
    if (v == 0) lse[li] = mx + __logf(e0 + e1);
  }
}

static AttnTcArgs window_args(const AttnTcArgs& a, int w) {
  AttnTcArgs x = a;
  constexpr int kWin = 256;                            // text keys per window
  if (w == 0) { x.kt0 = 0; x.Lk = a.L < kWin ? a.L : kWin; x.kbase = a.P; }
  else { x.P = 0; x.P8 = 0; x.kt0 = kWin; x.Lk = a.L - kWin; x.kbase = a.P + kWin; }
  x.L64 = (x.Lk + 63) / 64 * 64;
  x.N16 = (x.P8 + x.Lk + 15) / 16 * 16;
  return x;
}

bool attn_fwd_tc_windows_supported(const AttnTcArgs& a) {
  if (a.L <= 256 || a.L > 512 || a.P8 > 128) return false;
  return attn_fwd_tc_fits(window_args(a, 0)) && attn_fwd_tc_fits(window_args(a, 1));
}

size_t attn_fwd_tc_windows_workspace(int B, int L, int nh) {
  const size_t ctx1 = ((size_t)B * L * nh * 64 * sizeof(__nv_bfloat16) + 255) / 256 * 256;
  return ctx1 + (size_t)B * nh * L * sizeof(float);
}

int attn_fwd_tc_windows_launch(const AttnTcArgs& a, const AttnTcMaps& m, void* ctx, int64_t ld_ctx, float* lse,
                               void* workspace, cudaStream_t st) {
  const size_t ctx1_bytes = ((size_t)a.B * a.L * a.nh * 64 * sizeof(__nv_bfloat16) + 255) / 256 * 256;
  __nv_bfloat16* ctx1 = reinterpret_cast<__nv_bfloat16*>(workspace);
  float* lse1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ctx1_bytes);
  if (int rc = attn_fwd_tc_launch(window_args(a, 0), m, ctx, ld_ctx, lse, st)) return rc;
  if (int rc = attn_fwd_tc_launch(window_args(a, 1), m, ctx1, (int64_t)a.nh * 64, lse1, st)) return rc;
  const long long n = (long long)a.B * a.L * a.nh * 8;
  const int grid = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 16);
  attn_merge_windows_kernel<<<grid, 256, 0, st>>>((__nv_bfloat16*)ctx, ld_ctx, lse, ctx1, lse1, a.B, a.L, a.nh);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

// shape gate + tensor maps shared by forward and backward
int attn_tc_prepare(const void* qkv, int64_t ld_qkv, const void* kp, const void* vp, int P, const int64_t* key_mask,
                    int B, int L, int nh, float p_drop, uint64_t seed, AttnTcArgs* a, AttnTcMaps* m, bool* ok) {
  *ok = false;
  a->P = P; a->P8 = (P + 7) / 8 * 8; a->L = L; a->L64 = (L + 63) / 64 * 64;
  a->N16 = (a->P8 + L + 15) / 16 * 16;
  a->Lk = L; a->kt0 = 0; a->kbase = P;
  a->B = B; a->nh = nh; a->key_mask = reinterpret_cast<const long long*>(key_mask);
  a->scale = 0.125f;
  a->drop_thr = 0; a->drop_scale = 1.f; a->seed = seed; a->step = step_source();
  if (p_drop > 0.f) {
    double t = (double)p_drop * 4294967296.0;
    a->drop_thr = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    a->drop_scale = 1.f / (1.f - p_drop);
  }
  // (whether all keys of an item fit the resident-key kernels is the callers' question: attn_fwd_tc_fits,
  //  attn_bwd_*_supported; the tensor maps below serve every tcgen05 attention kernel)
  if (ld_qkv % 8 != 0 || (reinterpret_cast<uintptr_t>(qkv) & 15)) return 0;
  const uint64_t T = (uint64_t)B * L;
  const uint64_t W = 3ull * nh * 64;
  int rc = make_tmap_bf16_2d(&m->q, qkv, W, T, ld_qkv, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&m->kv, qkv, W, T, ld_qkv, 64, 64);
  if (rc) return rc;
  if (P > 0) {
    if ((reinterpret_cast<uintptr_t>(kp) & 15) || (reinterpret_cast<uintptr_t>(vp) & 15)) return 0;
    rc = make_tmap_bf16_2d(&m->kp, kp, 64, (uint64_t)B * nh * P, 64, 64, 8);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&m->vp, vp, 64, (uint64_t)B * nh * P, 64, 64, 8);
    if (rc) return rc;
  } else {
    m->kp = m->kv; m->vp = m->kv;
  }
  *ok = true;
  return 0;
}

}  // namespace mtvaf
