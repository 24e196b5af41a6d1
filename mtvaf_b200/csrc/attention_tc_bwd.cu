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
// keys of an item are resident.  Per item:
//   TMA  : {Q, K_p, K} and {dO, V_p, V} on two mbarriers (S can start before dO/V land); the loads of the
//          NEXT item are issued as soon as this item's last MMA has retired, i.e. they overlap the stores
//   MMA  : S -> TMEM cols [0,N16), dP -> TMEM cols [256,256+N16)          (both K-major operands)
//   SIMT : thread = (query row, quarter of the 8-key units): rowsum(dO o O) from a quarter row of dO (smem)
//          and O (global), exchanged through smem; S and dP from TMEM -> bf16 P and dS into shared memory in
//          the K-major SWIZZLE_128B layout [q][64-key chunk] (one 16-byte piece per 8-key unit)
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
  int n_chunks, kv_rows;
  size_t off_ds, off_p, off_q, off_do, off_k, off_v, off_mask, off_exch, off_bar, total;
};

__host__ __device__ inline AttnBwdSmem attn_bwd_layout(int P8, int L64, int N16) {
  AttnBwdSmem s;
  s.n_chunks = (N16 + 63) / 64;
  const int loaded = P8 + L64;
  s.kv_rows = loaded > N16 ? loaded : N16;
  size_t o = 0;
  s.off_ds = o; o += (size_t)s.n_chunks * 16384;      // dS chunks; a 128-key tile may read one chunk past
  s.off_p = o;  o += (size_t)s.n_chunks * 16384;      // the end (rows of the output that are never stored)
  s.off_q = o;  o += 16384;
  s.off_do = o; o += 16384;
  s.off_k = o;  o += (size_t)s.kv_rows * 128;
  s.off_v = o;  o += (size_t)s.kv_rows * 128;
  s.off_mask = o; o += (size_t)((N16 + 15) / 16) * 64;
  s.off_exch = o; o += 4 * 128 * sizeof(float);
  s.off_bar = o; o += 64;
  s.total = o + 1024;                                  // + alignment slack
  return s;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmKp, const __grid_constant__ CUtensorMap tmVp,
                   const __grid_constant__ CUtensorMap tmdO, AttnTcArgs a, const float* __restrict__ lse,
                   const __nv_bfloat16* __restrict__ ctx, long long ld_ctx, __nv_bfloat16* __restrict__ dqkv,
                   long long ld_dqkv, float* __restrict__ dkp, float* __restrict__ dvp) {
  constexpr int DP_COL = 256;                          // TMEM column of dP (S at 0)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const AttnBwdSmem lay = attn_bwd_layout(a.P8, a.L64, a.N16);
  uint8_t* sdS = smem + lay.off_ds;
  uint8_t* sP = smem + lay.off_p;
  uint8_t* sQ = smem + lay.off_q;
  uint8_t* sdO = smem + lay.off_do;
  uint8_t* sK = smem + lay.off_k;
  uint8_t* sV = smem + lay.off_v;
  float* sMask = reinterpret_cast<float*>(smem + lay.off_mask);           // additive mask * log2(e)
  float* sExch = reinterpret_cast<float*>(smem + lay.off_exch);           // [4][128] partial rowsum(dO o O)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.off_bar);       // q/k, dO/v, s/dp, grads
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;         // TMEM lane group / quarter of the columns
  const int row = quad * 32 + lane;                    // query row == TMEM lane
  const int H = a.nh * 64;
  const int loaded_rows = a.P8 + a.L64;
  const int n_items = a.B * a.nh;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmKV); prefetch_tmap(&tmdO);
    if (a.P8 > 0) { prefetch_tmap(&tmKp); prefetch_tmap(&tmVp); }
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  // K/V rows the TMA boxes do not cover but the MMAs read: zero them once (0 x garbage could be NaN)
  for (int i = loaded_rows * 8 + tid; i < a.N16 * 8; i += kBwdThreads) {
    *reinterpret_cast<uint4*>(sK + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(sV + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto issue_loads = [&](int item) {                   // one thread
    const int b = item / a.nh, h = item - b * a.nh;
    const uint32_t bytes = 16384u + (uint32_t)loaded_rows * 128u;
    mbar_arrive_expect_tx(&bars[0], bytes);
    tma_load_2d(sQ, &tmQ, &bars[0], h * 64, b * a.L);
    for (int r = 0; r < a.P8; r += 8) tma_load_2d(sK + r * 128, &tmKp, &bars[0], 0, (b * a.nh + h) * a.P + r);
    for (int r = 0; r < a.L64; r += 64)
      tma_load_2d(sK + (a.P8 + r) * 128, &tmKV, &bars[0], H + h * 64, b * a.L + r);
    mbar_arrive_expect_tx(&bars[1], bytes);
    tma_load_2d(sdO, &tmdO, &bars[1], h * 64, b * a.L);
    for (int r = 0; r < a.P8; r += 8) tma_load_2d(sV + r * 128, &tmVp, &bars[1], 0, (b * a.nh + h) * a.P + r);
    for (int r = 0; r < a.L64; r += 64)
      tma_load_2d(sV + (a.P8 + r) * 128, &tmKV, &bars[1], 2 * H + h * 64, b * a.L + r);
  };
  if (tid == 0 && (int)blockIdx.x < n_items) issue_loads(blockIdx.x);

  const bool row_ok = row < a.L;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  const float sc2 = a.scale * kLog2e;
  const int row8 = row & 7;
  const uint32_t prow_off = (row >> 3) * 1024 + row8 * 128;
  const int units = a.N16 >> 3;                        // 8-key units (one 16-byte smem piece each)
  const int n_tiles = (a.N16 + 127) / 128;             // 128-key tiles of dK / dV
  constexpr int DQ_COL = 0, DK_COL = 64, DV_COL = DP_COL;   // dK tiles at 64, 128 ; dV tiles at 256, 320
  const int dcol = part * 16;                          // this thread's 16 of the 64 head-dim columns

  uint32_t ph = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ph ^= 1) {
    const int b = item / a.nh, h = item - b * a.nh;
    // ---- key validity / additive mask (x log2 e) in smem-key numbering: 0 visible, -10000 padded text key,
    //      -inf = no such key (prefix padding, rows past L)
    for (int k = tid; k < a.N16; k += kBwdThreads) {
      float m;
      if (k < a.P8) m = (k < a.P) ? 0.f : -INFINITY;
      else {
        const int t = k - a.P8;
        m = (t < a.L) ? (a.key_mask[(long long)b * a.L + t] != 0 ? 0.f : -10000.0f * kLog2e) : -INFINITY;
      }
      sMask[k] = m;
    }
    // ---- per-row scalars and this thread's quarter row of O (global, 32 B)
    const float lse2 = row_ok ? lse[((long long)b * a.nh + h) * a.L + row] * kLog2e : INFINITY;
    uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
    if (row_ok) {
      const uint4* po = reinterpret_cast<const uint4*>(ctx + ((long long)b * a.L + row) * ld_ctx + h * 64 + dcol);
      o0 = po[0];
      o1 = po[1];
    }
    if (tid == 0) {
      // ---- S = Q K^T -> cols [0,N16) ; dP = dO V^T -> cols [256, 256+N16)
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), adO = smem_u32(sdO), aV = smem_u32(sV);
      const uint32_t idesc = make_idesc_bf16(128, a.N16, false, false);
      mbar_wait(&bars[0], ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tmem_base, make_smem_desc_sw128(aQ + k * 32, 16, 1024),
                    make_smem_desc_sw128(aK + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
      mbar_wait(&bars[1], ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tmem_base + DP_COL, make_smem_desc_sw128(adO + k * 32, 16, 1024),
                    make_smem_desc_sw128(aV + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
      umma_commit(&bars[2]);
    }
    __syncwarp();
    // ---- D_q = rowsum(dO o O): this thread's 16 columns (two 16-byte pieces of the swizzled dO row)
    mbar_wait(&bars[1], ph);
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
        a.drop_thr ? attn_drop_rowkey(a.seed, ((unsigned long long)b * a.nh + h) * a.L + row) : 0u;

    mbar_wait(&bars[2], ph);
    __syncwarp();
    tc_fence_after();

    // ---- P and dS for this thread's (row, every 4th 8-key unit)
    for (int u = part; u < units; u += 4) {
      const int c = u << 3;
      uint32_t rs[8], rd[8];
      tmem_ld_32x32b_x8(t_row + c, rs);
      tmem_ld_32x32b_x8(t_row + DP_COL + c, rd);
      const float4 m0 = *reinterpret_cast<const float4*>(sMask + c);
      const float4 m1 = *reinterpret_cast<const float4*>(sMask + c + 4);
      const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      bool keep[8];
      if (a.drop_thr) attn_drop_keep8(rowkey, c < a.P8 ? c : a.P + (c - a.P8), a.drop_thr, keep);
      tmem_ld_wait();
      float p[8], ds[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // masked / absent keys and rows past L give exp2(-inf) = 0 (S and dP are finite: padded K/V rows are 0)
        const float pj = exp2f(fmaf(__uint_as_float(rs[j]), sc2, mk[j] - lse2));
        float dp = __uint_as_float(rd[j]);
        float pk = pj;
        if (a.drop_thr) {
          dp = keep[j] ? dp * a.drop_scale : 0.f;
          pk = keep[j] ? pj * a.drop_scale : 0.f;
        }
        ds[j] = (pj * a.scale) * (dp - dsum);
        p[j] = pk;
      }
      const uint32_t off = (u >> 3) * 16384 + prow_off + (((u & 7) ^ row8) << 4);
      uint4 w;
      w.x = pack_bf16x2(p[0], p[1]); w.y = pack_bf16x2(p[2], p[3]);
      w.z = pack_bf16x2(p[4], p[5]); w.w = pack_bf16x2(p[6], p[7]);
      *reinterpret_cast<uint4*>(sP + off) = w;
      w.x = pack_bf16x2(ds[0], ds[1]); w.y = pack_bf16x2(ds[2], ds[3]);
      w.z = pack_bf16x2(ds[4], ds[5]); w.w = pack_bf16x2(ds[6], ds[7]);
      *reinterpret_cast<uint4*>(sdS + off) = w;
    }
    // key columns [N16, 64*n_chunks) of the last chunk are read by the key-tile MMAs as rows that are never
    // stored; they need no initialisation (TMEM lanes are independent).
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (tid == 0) {
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
      umma_commit(&bars[3]);
    }
    __syncwarp();
    mbar_wait(&bars[3], ph);
    __syncwarp();
    tc_fence_after();
    // every operand tile of this item has been consumed: fetch the next item's while the gradients drain
    if (tid == 0 && item + (int)gridDim.x < n_items) issue_loads(item + gridDim.x);
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
  return attn_bwd_layout(a.P8, a.L64, a.N16).total <= 227 * 1024;
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
  const AttnBwdSmem lay = attn_bwd_layout(a.P8, a.L64, a.N16);
  static bool set = false;
  if (!set) {
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  const int n_items = a.B * a.nh;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  attn_bwd_tc_kernel<<<grid, kBwdThreads, lay.total, st>>>(m.q, m.kv, m.kp, m.vp, tmdO, a, lse,
                                                          (const __nv_bfloat16*)ctx, ld_ctx, (__nv_bfloat16*)dqkv,
                                                          ld_dqkv, dkp, dvp);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtvaf
