// Backward of the prefix ("fusion") self-attention on the tcgen05 tensor cores (bf16 in, fp32 in TMEM).
//
// Autograd of RobertaSelfAttention.forward models/modeling_roberta.py:218-278 in the flash-attention
// form (saved log-sum-exp, probabilities recomputed, never stored):
//   S  = Q K^T                       P  = exp(S/sqrt(d) + mask - lse)       (dropout mask M regenerated)
//   dP = dO V^T                      dS = P o (M dP/(1-p) - rowsum(dO o O)) / sqrt(d)
//   dQ = dS K        dK = dS^T Q        dV = (M P/(1-p))^T dO
// Keys are numbered prefix rows first (padded to a multiple of 8) then text rows -- the torch.cat of
// :221-222 -- so the gradient of the visual prefix (dK_p, dV_p -> get_visual_prompt,
// models/bert_model.py:566-587) falls out of the same two MMAs as the text dK/dV.
//
// PERSISTENT: one CTA (512 threads) per SM walks the (batch, head) items; all L <= 128 queries x all P+L
// keys of an item are resident.  The kernel is bound by the latency chain load -> MMA -> SIMT -> MMA -> store and
// by ~130 KB of HBM traffic per item, so everything that can run ahead does:
//   TMA  : {Q, K_p, K} are DOUBLE-buffered (when shared memory allows): the next item's tiles are requested at the
//          top of the current item; V is requested as soon as dP = dO V^T has retired, dO as soon as the last MMA
//          has retired (behind the stores); each has its own mbarrier, S starts as soon as Q/K are there
//   regs : the next item's key-mask value, log-sum-exp and quarter row of O are fetched one item ahead
//   MMA  : S -> TMEM cols [0,N16), dP -> TMEM cols [256,256+N16)          (both K-major operands)
//   SIMT : thread = (query row, quarter of the 8-key units): rowsum(dO o O) from a quarter row of dO (smem)
//          and O (registers), exchanged through smem; S and dP from TMEM -> bf16 P and dS into shared memory in
//          the K-major SWIZZLE_128B layout [q][64-key chunk] (one 16-byte piece per 8-key unit); the dropout
//          scale rides in the exponent, the mask is one hash per PAIR of keys
//   MMA  : dQ = dS K (A = dS K-major, B = K MN-major); dK = dS^T Q and dV = P^T dO per 128-key tile
//          (A = the SAME dS / P buffers read as MN-major operands, B = Q / dO MN-major);
//          accumulators alias the S / dP columns
//   store: dQ, dK, dV -> d(qkv) [T, 3H] bf16 (32 B per thread per row); prefix rows -> fp32 dK_p / dV_p.
// TMEM, barriers and the zero padding of K/V are set up once per CTA.
#include "attention_tc.cuh"

namespace mtvaf {
using namespace ptx;

constexpr int kBwdThreads = 512;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnBwdSmem {
  int n_chunks, kv_rows, nbuf;
  size_t off_ds, off_p, off_q, off_do, off_k, off_v, off_mask, off_exch, off_bar, total;
  size_t q_stride, k_stride;
};

__host__ __device__ inline AttnBwdSmem attn_bwd_layout(int P8, int L64, int N16, int nbuf) {
  AttnBwdSmem s;
  s.n_chunks = (N16 + 63) / 64;
  s.nbuf = nbuf;
  const int loaded = P8 + L64;
  s.kv_rows = loaded > N16 ? loaded : N16;
  s.q_stride = 16384;
  s.k_stride = (((size_t)s.kv_rows * 128) + 1023) / 1024 * 1024;
  size_t o = 0;
  s.off_ds = o; o += (size_t)s.n_chunks * 16384;      // dS chunks; a 128-key tile may read one chunk past
  s.off_p = o;  o += (size_t)s.n_chunks * 16384;      // the end (rows of the output that are never stored)
  s.off_q = o;  o += s.q_stride * nbuf;
  s.off_do = o; o += 16384;
  s.off_k = o;  o += s.k_stride * nbuf;
  s.off_v = o;  o += s.k_stride;
  s.off_mask = o; o += (size_t)((N16 + 15) / 16) * 64;
  s.off_exch = o; o += 4 * 128 * sizeof(float);
  s.off_bar = o; o += 64;
  s.total = o + 1024;                                  // + alignment slack
  return s;
}

// 2 (double-buffered Q/K) when it fits in shared memory, else 1, else 0 (shape not supported)
static int attn_bwd_pick_nbuf(const AttnTcArgs& a) {
  if (attn_bwd_layout(a.P8, a.L64, a.N16, 2).total <= 227 * 1024) return 2;
  if (attn_bwd_layout(a.P8, a.L64, a.N16, 1).total <= 227 * 1024) return 1;
  return 0;
}

__device__ __forceinline__ float bwd_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmKp, const __grid_constant__ CUtensorMap tmVp,
                   const __grid_constant__ CUtensorMap tmdO, AttnTcArgs a, int nbuf, const float* __restrict__ lse,
                   const __nv_bfloat16* __restrict__ ctx, long long ld_ctx, __nv_bfloat16* __restrict__ dqkv,
                   long long ld_dqkv, float* __restrict__ dkp, float* __restrict__ dvp) {
  constexpr int DP_COL = 256;                          // TMEM column of dP (S at 0)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const AttnBwdSmem lay = attn_bwd_layout(a.P8, a.L64, a.N16, nbuf);
  uint8_t* sdS = smem + lay.off_ds;
  uint8_t* sP = smem + lay.off_p;
  uint8_t* sQ0 = smem + lay.off_q;
  uint8_t* sdO = smem + lay.off_do;
  uint8_t* sK0 = smem + lay.off_k;
  uint8_t* sV = smem + lay.off_v;
  float* sMask = reinterpret_cast<float*>(smem + lay.off_mask);           // additive mask * log2(e)
  float* sExch = reinterpret_cast<float*>(smem + lay.off_exch);           // [4][128] partial rowsum(dO o O)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.off_bar);       // qk[0], qk[1], dO, V, s/dp, grads
  uint64_t* bar_qk = bars;
  uint64_t* bar_do = bars + 2;
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_s = bars + 4;
  uint64_t* bar_g = bars + 5;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;         // TMEM lane group / quarter of the columns
  const int row = quad * 32 + lane;                    // query row == TMEM lane
  const int H = a.nh * 64;
  const int loaded_rows = a.P8 + a.L64;
  const int n_items = a.B * a.nh;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmKV); prefetch_tmap(&tmdO);
    if (a.P8 > 0) { prefetch_tmap(&tmKp); prefetch_tmap(&tmVp); }
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  // K/V rows the TMA boxes do not cover but the MMAs read: zero them once (0 x garbage could be NaN)
  for (int i = loaded_rows * 8 + tid; i < a.N16 * 8; i += kBwdThreads) {
    for (int s = 0; s < nbuf; ++s) *reinterpret_cast<uint4*>(sK0 + s * lay.k_stride + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(sV + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t aMask = smem_u32(sMask), aPs = smem_u32(sP), adSs = smem_u32(sdS);   // explicit shared-space accesses

  const uint32_t kv_bytes = (uint32_t)loaded_rows * 128u;
  auto load_qk = [&](int item, int buf) {              // one thread
    const int b = item / a.nh, h = item - b * a.nh;
    uint8_t* q = sQ0 + buf * lay.q_stride;
    uint8_t* k = sK0 + buf * lay.k_stride;
    mbar_arrive_expect_tx(&bar_qk[buf], 16384u + kv_bytes);
    tma_load_2d(q, &tmQ, &bar_qk[buf], h * 64, b * a.L);
    for (int r = 0; r < a.P8; r += 8) tma_load_2d(k + r * 128, &tmKp, &bar_qk[buf], 0, (b * a.nh + h) * a.P + r);
    for (int r = 0; r < a.L64; r += 64)
      tma_load_2d(k + (a.P8 + r) * 128, &tmKV, &bar_qk[buf], H + h * 64, b * a.L + r);
  };
  auto load_v = [&](int item) {
    const int b = item / a.nh, h = item - b * a.nh;
    mbar_arrive_expect_tx(bar_v, kv_bytes);
    for (int r = 0; r < a.P8; r += 8) tma_load_2d(sV + r * 128, &tmVp, bar_v, 0, (b * a.nh + h) * a.P + r);
    for (int r = 0; r < a.L64; r += 64)
      tma_load_2d(sV + (a.P8 + r) * 128, &tmKV, bar_v, 2 * H + h * 64, b * a.L + r);
  };
  auto load_do = [&](int item) {
    const int b = item / a.nh, h = item - b * a.nh;
    mbar_arrive_expect_tx(bar_do, 16384u);
    tma_load_2d(sdO, &tmdO, bar_do, h * 64, b * a.L);
  };

  const bool row_ok = row < a.L;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  const float sc2 = a.scale * kLog2e;
  const int row8 = row & 7;
  const uint32_t prow_off = (row >> 3) * 1024 + row8 * 128;
  const int units = a.N16 >> 3;                        // 8-key units (one 16-byte smem piece each)
  const int n_tiles = (a.N16 + 127) / 128;             // 128-key tiles of dK / dV
  constexpr int DQ_COL = 0, DK_COL = 64, DV_COL = DP_COL;   // dK tiles at 64, 128 ; dV tiles at 256, 320
  const int dcol = part * 16;                          // this thread's 16 of the 64 head-dim columns
  const float log2_ds = a.drop_thr ? log2f(a.drop_scale) : 0.f;
  const float ds_coef = a.scale / a.drop_scale;        // dS = P' * (scale / drop_scale) * (dP' - D)

  // per-item values fetched one item ahead: key-mask entry of key `tid`, log-sum-exp, quarter row of O
  auto fetch_mask = [&](int item) -> float {
    if (tid >= a.N16) return 0.f;
    const int b = item / a.nh;
    if (tid < a.P8) return (tid < a.P) ? 0.f : -INFINITY;
    const int t = tid - a.P8;
    return (t < a.L) ? (a.key_mask[(long long)b * a.L + t] != 0 ? 0.f : -10000.0f * kLog2e) : -INFINITY;
  };
  auto fetch_lse = [&](int item) -> float {
    const int b = item / a.nh, h = item - b * a.nh;
    // +inf for rows past L: P = exp2(-inf) = 0
    return row_ok ? lse[((long long)b * a.nh + h) * a.L + row] * kLog2e - log2_ds : INFINITY;
  };
  auto fetch_o = [&](int item, uint4& o0, uint4& o1) {
    o0 = make_uint4(0, 0, 0, 0);
    o1 = o0;
    if (row_ok) {
      const int b = item / a.nh, h = item - b * a.nh;
      const uint4* po = reinterpret_cast<const uint4*>(ctx + ((long long)b * a.L + row) * ld_ctx + h * 64 + dcol);
      o0 = po[0];
      o1 = po[1];
    }
  };

  const int first = blockIdx.x;
  float m_next = 0.f, lse_next = 0.f;
  uint4 o0n = make_uint4(0, 0, 0, 0), o1n = o0n;
  if (first < n_items) {
    __syncwarp();                                    // elect.sync needs the whole warp
    if (warp == 0 && elect_one()) { load_qk(first, 0); load_do(first); load_v(first); }
    m_next = fetch_mask(first);
    lse_next = fetch_lse(first);
    fetch_o(first, o0n, o1n);
  }

  int il = 0;                                          // local iteration count
  for (int item = first; item < n_items; item += gridDim.x, ++il) {
    const int b = item / a.nh, h = item - b * a.nh;
    const int buf = nbuf == 2 ? (il & 1) : 0;
    const uint32_t ph = il & 1;
    const uint32_t ph_qk = nbuf == 2 ? ((il >> 1) & 1) : ph;
    const int next = item + gridDim.x;
    uint8_t* sQ = sQ0 + buf * lay.q_stride;
    uint8_t* sK = sK0 + buf * lay.k_stride;
    // ---- this item's prefetched scalars; key validity / additive mask (x log2 e) in smem-key numbering:
    //      0 visible, -10000 padded text key, -inf = no such key (prefix padding, rows past L)
    if (tid < a.N16) sMask[tid] = m_next;
    const float lse2 = lse_next;
    const uint4 o0 = o0n, o1 = o1n;
    __syncwarp();                                    // elect.sync needs the whole warp
    if (warp == 0 && elect_one()) {
      // the other Q/K buffer was last read by the previous item's MMAs (retired): request the next item's tiles
      if (nbuf == 2 && next < n_items) load_qk(next, buf ^ 1);
      // ---- S = Q K^T -> cols [0,N16) ; dP = dO V^T -> cols [256, 256+N16)
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), adO = smem_u32(sdO), aV = smem_u32(sV);
      const uint32_t idesc = make_idesc_bf16(128, a.N16, false, false);
      mbar_wait(&bar_qk[buf], ph_qk);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tmem_base, make_smem_desc_sw128(aQ + k * 32, 16, 1024),
                    make_smem_desc_sw128(aK + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
      mbar_wait(bar_do, ph);
      mbar_wait(bar_v, ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tmem_base + DP_COL, make_smem_desc_sw128(adO + k * 32, 16, 1024),
                    make_smem_desc_sw128(aV + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
    // ---- D_q = rowsum(dO o O): this thread's 16 columns (two 16-byte pieces of the swizzled dO row)
    mbar_wait(bar_do, ph);
    {
      const uint8_t* pd = sdO + prow_off;
      const uint4 d0 = *reinterpret_cast<const uint4*>(pd + (((part * 2) ^ row8) << 4));
      const uint4 d1 = *reinterpret_cast<const uint4*>(pd + (((part * 2 + 1) ^ row8) << 4));
      float acc = 0.f;
      const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      const uint32_t ow[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 x = unpack_bf16x2(dw[j]), y = unpack_bf16x2(ow[j]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
      sExch[part * 128 + row] = acc;
    }
    __syncthreads();                                   // publishes sMask and sExch
    const float dsum = (sExch[row] + sExch[128 + row]) + (sExch[256 + row] + sExch[384 + row]);
    const uint32_t rowkey =
        a.drop_thr ? attn_drop_rowkey(step_seed(a.seed, a.step), ((unsigned long long)b * a.nh + h) * a.L + row) : 0u;
    // the next item's scalars travel while this item's softmax runs
    if (next < n_items) {
      m_next = fetch_mask(next);
      lse_next = fetch_lse(next);
      fetch_o(next, o0n, o1n);
    }

    mbar_wait(bar_s, ph);
    __syncwarp();
    tc_fence_after();
    __syncwarp();                                    // elect.sync needs the whole warp
    if (warp == 0 && next < n_items && elect_one()) load_v(next);      // dP has retired: V is free

    // ---- P and dS for this thread's (row, every 4th 8-key unit).  P' = P / (1-p) is born scaled (the dropout scale
    // rides in the exponent: lse2 already carries -log2(scale)); dropped keys are zeroed in P' and in dP.
    for (int u = part; u < units; u += 4) {
      const int c = u << 3;
      uint32_t rs[8], rd[8];
      tmem_ld_32x32b_x8(t_row + c, rs);
      tmem_ld_32x32b_x8(t_row + DP_COL + c, rd);
      const float4 m0 = lds_f4(aMask + (c) * 4);
      const float4 m1 = lds_f4(aMask + (c) * 4 + 16);
      const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      tmem_ld_wait();
      float p[8], dp[8], ds[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // masked / absent keys and rows past L give exp2(-inf) = 0 (S and dP are finite: padded K/V rows are 0)
        p[j] = bwd_ex2(fmaf(__uint_as_float(rs[j]), sc2, mk[j] - lse2));
        dp[j] = __uint_as_float(rd[j]) * a.drop_scale;
        ds[j] = p[j] * ds_coef;
      }
      if (a.drop_thr) attn_drop_apply8(rowkey, c < a.P8 ? c : a.P + (c - a.P8), a.drop_thr, p, dp);
#pragma unroll
      for (int j = 0; j < 8; ++j) ds[j] *= dp[j] - dsum;
      const uint32_t off = (u >> 3) * 16384 + prow_off + (((u & 7) ^ row8) << 4);
      uint4 w;
      w.x = pack_bf16x2(p[0], p[1]); w.y = pack_bf16x2(p[2], p[3]);
      w.z = pack_bf16x2(p[4], p[5]); w.w = pack_bf16x2(p[6], p[7]);
      sts_u4(aPs + off, w);
      w.x = pack_bf16x2(ds[0], ds[1]); w.y = pack_bf16x2(ds[2], ds[3]);
      w.z = pack_bf16x2(ds[4], ds[5]); w.w = pack_bf16x2(ds[6], ds[7]);
      sts_u4(adSs + off, w);
    }
    // key columns [N16, 64*n_chunks) of the last chunk are read by the key-tile MMAs as rows that are never
    // stored; they need no initialisation (TMEM lanes are independent).
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    __syncwarp();                                    // elect.sync needs the whole warp
    if (warp == 0 && elect_one()) {
      const uint32_t adS = smem_u32(sdS), aP = smem_u32(sP), aQ = smem_u32(sQ), adO = smem_u32(sdO),
                     aK = smem_u32(sK);
      // dQ[q, d] = sum_key dS[q,key] K[key,d]
      {
        const uint32_t idesc = make_idesc_bf16(128, 64, false, true);
        const int ksteps = a.N16 / 16;
        for (int j = 0; j < ksteps; ++j)
          umma_f16_ss(tmem_base + DQ_COL, make_smem_desc_sw128(adS + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                      make_smem_desc_sw128(aK + j * 2048, 8192, 1024), idesc, j > 0 ? 1u : 0u);
      }
      // dK[key, d] = sum_q dS[q,key] Q[q,d] ; dV[key, d] = sum_q P[q,key] dO[q,d]   (K dimension = 128 queries)
      const uint32_t idesc_t = make_idesc_bf16(128, 64, true, true);
      for (int t = 0; t < n_tiles; ++t) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_f16_ss(tmem_base + DK_COL + t * 64, make_smem_desc_sw128(adS + t * 32768 + j * 2048, 16384, 1024),
                      make_smem_desc_sw128(aQ + j * 2048, 8192, 1024), idesc_t, j > 0 ? 1u : 0u);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_f16_ss(tmem_base + DV_COL + t * 64, make_smem_desc_sw128(aP + t * 32768 + j * 2048, 16384, 1024),
                      make_smem_desc_sw128(adO + j * 2048, 8192, 1024), idesc_t, j > 0 ? 1u : 0u);
      }
      umma_commit(bar_g);
    }
    __syncwarp();
    mbar_wait(bar_g, ph);
    __syncwarp();
    tc_fence_after();
    // every operand tile of this item has been consumed: fetch the next item's while the gradients drain
    __syncwarp();                                    // elect.sync needs the whole warp
    if (warp == 0 && next < n_items && elect_one()) {
      load_do(next);
      if (nbuf == 1) load_qk(next, 0);
    }
    __syncwarp();

    // ---- stores: this thread owns 16 of the 64 head-dim columns of its row (TMEM loads are warp-collective:
    // all lanes issue them, only the global stores are predicated)
    {
      uint32_t r[16];
      tmem_ld_32x32b_x16(t_row + DQ_COL + dcol, r);
      tmem_ld_wait();
      if (row_ok) {
        uint4* o = reinterpret_cast<uint4*>(dqkv + ((long long)b * a.L + row) * ld_dqkv + h * 64 + dcol);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
          w.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
          w.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
          w.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
          o[v] = w;
        }
      }
    }
    for (int t = 0; t < n_tiles; ++t) {
      const int ks = t * 128 + row;                          // smem key number of this lane
      const bool is_prefix = ks < a.P;
      const int tx = ks - a.P8;
      const bool is_text = ks >= a.P8 && tx < a.L;
#pragma unroll
      for (int which = 0; which < 2; ++which) {              // 0: dK, 1: dV
        uint32_t r[16];
        __syncwarp();
        tmem_ld_32x32b_x16(t_row + (which ? DV_COL : DK_COL) + t * 64 + dcol, r);
        tmem_ld_wait();
        if (is_text) {
          uint4* o = reinterpret_cast<uint4*>(dqkv + ((long long)b * a.L + tx) * ld_dqkv + (which + 1) * H + h * 64 +
                                              dcol);
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
            w.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
            w.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
            w.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
            o[v] = w;
          }
        } else if (is_prefix) {
          float* o = (which ? dvp : dkp);
          if (o) {
            o += (((long long)b * a.nh + h) * a.P + ks) * 64 + dcol;
#pragma unroll
            for (int v = 0; v < 4; ++v)
              *reinterpret_cast<float4*>(o + v * 4) =
                  make_float4(__uint_as_float(r[v * 4 + 0]), __uint_as_float(r[v * 4 + 1]),
                              __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
          }
        }
      }
    }
    // all TMEM reads of this item done before the next item's S / dP MMAs overwrite the columns;
    // also orders this item's sMask / sExch reads before the next item's writes
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// shape gate of the backward kernel (on top of attn_tc_prepare's)
bool attn_bwd_tc_supported(const AttnTcArgs& a) {
  if (a.L > 128 || a.N16 > 256) return false;
  return attn_bwd_pick_nbuf(a) > 0;
}

int attn_bwd_tc_launch(const AttnTcArgs& a, const AttnTcMaps& m, const void* dctx, int64_t ld_dctx, const void* ctx,
                       int64_t ld_ctx, const float* lse, void* dqkv, int64_t ld_dqkv, float* dkp, float* dvp,
                       cudaStream_t st) {
  MTVAF_REQUIRE(ld_dqkv % 8 == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0,
                "attention_bwd(tc): dqkv must be 16-byte aligned with ld %% 8 == 0");
  MTVAF_REQUIRE(ld_ctx % 8 == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0,
                "attention_bwd(tc): ctx must be 16-byte aligned with ld %% 8 == 0");
  CUtensorMap tmdO;
  const uint64_t T = (uint64_t)a.B * a.L;
  const uint64_t H = (uint64_t)a.nh * 64;
  int rc = make_tmap_bf16_2d(&tmdO, dctx, H, T, ld_dctx, 64, 128);
  if (rc) return rc;
  const int nbuf = attn_bwd_pick_nbuf(a);
  MTVAF_REQUIRE(nbuf > 0, "attention_bwd(tc): shape does not fit in shared memory");
  const AttnBwdSmem lay = attn_bwd_layout(a.P8, a.L64, a.N16, nbuf);
  static bool set = false;
  if (!set) {
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  const int n_items = a.B * a.nh;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  attn_bwd_tc_kernel<<<grid, kBwdThreads, lay.total, st>>>(m.q, m.kv, m.kp, m.vp, tmdO, a, nbuf, lse,
                                                          (const __nv_bfloat16*)ctx, ld_ctx, (__nv_bfloat16*)dqkv,
                                                          ld_dqkv, dkp, dvp);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtvaf
