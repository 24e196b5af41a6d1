// Backward of the prefix ("fusion") self-attention on the tcgen05 tensor cores (bf16 in, fp32 in TMEM).
//
// Autograd of RobertaSelfAttention.forward models/modeling_roberta.py:218-278 in the flash-attention
// form (saved log-sum-exp, probabilities recomputed, never stored):
//   S  = Q K^T                       P  = exp(S/sqrt(d) + mask - lse)       (dropout mask M regenerated)
//   dP = dO V^T                      dS = P o (M dP/(1-p) - rowsum(dO o O)) / sqrt(d)
//   dQ = dS K        dK = dS^T Q        dV = (M P/(1-p))^T dO
// One CTA (256 threads) per (batch, head), all L <= 128 queries x all P+L keys; keys are numbered
// prefix rows first (padded to a multiple of 8) then text rows -- the torch.cat of :221-222 -- so the
// gradient of the visual prefix (dK_p, dV_p -> get_visual_prompt, models/bert_model.py:566-587) falls
// out of the same two MMAs as the text dK/dV.
//   TMA  : Q, dO, O tiles (128-row boxes), K_p/V_p (8-row boxes), K/V (64-row boxes), SWIZZLE_128B
//   MMA  : S -> TMEM cols [0,N16), dP -> TMEM cols [256,256+N16)          (both K-major operands)
//   SIMT : thread = (query row, half of the key columns): reads S and dP from TMEM, writes bf16 P and dS
//          into shared memory in the K-major SWIZZLE_128B layout [q][64-key chunk]
//   MMA  : dQ = dS K (A = dS K-major, B = K MN-major); dK = dS^T Q and dV = P^T dO per 128-key tile
//          (A = the SAME dS / P buffers read as MN-major operands, B = Q / dO MN-major);
//          accumulators alias the S / dP columns
//   store: dQ, dK, dV -> d(qkv) [T, 3H] bf16; prefix rows -> fp32 dK_p / dV_p.
#include "attention_tc.cuh"

namespace mtvaf {
using namespace ptx;

constexpr int kBwdThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnBwdSmem {
  int n_chunks, kv_rows;
  size_t off_ds, off_p, off_q, off_do, off_o, off_k, off_v, off_mask, off_bar, total;
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
  s.off_o = o;  o += 16384;
  s.off_k = o;  o += (size_t)s.kv_rows * 128;
  s.off_v = o;  o += (size_t)s.kv_rows * 128;
  s.off_mask = o; o += (size_t)((N16 + 15) / 16) * 64;
  s.off_bar = o; o += 64;
  s.total = o + 1024;                                  // + alignment slack
  return s;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmKp, const __grid_constant__ CUtensorMap tmVp,
                   const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmO,
                   AttnTcArgs a, const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv,
                   long long ld_dqkv, float* __restrict__ dkp, float* __restrict__ dvp) {
  constexpr int DP_COL = 256;                          // TMEM column of dP (S at 0)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const AttnBwdSmem lay = attn_bwd_layout(a.P8, a.L64, a.N16);
  uint8_t* sdS = smem + lay.off_ds;
  uint8_t* sP = smem + lay.off_p;
  uint8_t* sQ = smem + lay.off_q;
  uint8_t* sdO = smem + lay.off_do;
  uint8_t* sO = smem + lay.off_o;
  uint8_t* sK = smem + lay.off_k;
  uint8_t* sV = smem + lay.off_v;
  float* sMask = reinterpret_cast<float*>(smem + lay.off_mask);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.off_bar);       // load, s/dp, grads
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;                    // query row == TMEM lane
  const int b = blockIdx.z, h = blockIdx.y;
  const int H = a.nh * 64;
  const int loaded_rows = a.P8 + a.L64;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmKV); prefetch_tmap(&tmdO); prefetch_tmap(&tmO);
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  // key validity / additive mask in smem-key numbering: 0 visible, -10000 padded text key, -inf = no such key
  for (int k = tid; k < a.N16; k += kBwdThreads) {
    float m;
    if (k < a.P8) m = (k < a.P) ? 0.f : -INFINITY;
    else {
      const int t = k - a.P8;
      m = (t < a.L) ? (a.key_mask[(long long)b * a.L + t] != 0 ? 0.f : -10000.0f) : -INFINITY;
    }
    sMask[k] = m;
  }
  // K/V rows the TMA boxes do not cover but the MMAs read: zero them (0 x garbage could be NaN)
  for (int i = loaded_rows * 8 + tid; i < a.N16 * 8; i += kBwdThreads) {
    *reinterpret_cast<uint4*>(sK + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(sV + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (tid == 0) {
    const uint32_t bytes = 3u * 16384u + 2u * loaded_rows * 128u;
    mbar_arrive_expect_tx(&bars[0], bytes);
    tma_load_2d(sQ, &tmQ, &bars[0], h * 64, b * a.L);
    tma_load_2d(sdO, &tmdO, &bars[0], h * 64, b * a.L);
    tma_load_2d(sO, &tmO, &bars[0], h * 64, b * a.L);
    for (int r = 0; r < a.P8; r += 8) {
      tma_load_2d(sK + r * 128, &tmKp, &bars[0], 0, (b * a.nh + h) * a.P + r);
      tma_load_2d(sV + r * 128, &tmVp, &bars[0], 0, (b * a.nh + h) * a.P + r);
    }
    for (int r = 0; r < a.L64; r += 64) {
      tma_load_2d(sK + (a.P8 + r) * 128, &tmKV, &bars[0], H + h * 64, b * a.L + r);
      tma_load_2d(sV + (a.P8 + r) * 128, &tmKV, &bars[0], 2 * H + h * 64, b * a.L + r);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    // ---- S = Q K^T -> cols [0,N16) ; dP = dO V^T -> cols [256, 256+N16)
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), adO = smem_u32(sdO), aV = smem_u32(sV);
    const uint32_t idesc = make_idesc_bf16(128, a.N16, false, false);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_f16_ss(tmem_base, make_smem_desc_sw128(aQ + k * 32, 16, 1024),
                  make_smem_desc_sw128(aK + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_f16_ss(tmem_base + DP_COL, make_smem_desc_sw128(adO + k * 32, 16, 1024),
                  make_smem_desc_sw128(aV + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  __syncwarp();
  // every thread needs the operand tiles in smem too (dO, O for the row sums)
  mbar_wait(&bars[0], 0);

  // ---- D_q = rowsum(dO o O)  (each thread reads its own row; the 128B swizzle makes this conflict-free)
  const bool row_ok = row < a.L;
  float dsum = 0.f;
  {
    const uint8_t* pd = sdO + (row >> 3) * 1024 + (row & 7) * 128;
    const uint8_t* po = sO + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int sw = ((c ^ (row & 7)) << 4);
      const uint4 ud = *reinterpret_cast<const uint4*>(pd + sw);
      const uint4 uo = *reinterpret_cast<const uint4*>(po + sw);
      const float2 d0 = unpack_bf16x2(ud.x), d1 = unpack_bf16x2(ud.y), d2 = unpack_bf16x2(ud.z), d3 = unpack_bf16x2(ud.w);
      const float2 o0 = unpack_bf16x2(uo.x), o1 = unpack_bf16x2(uo.y), o2 = unpack_bf16x2(uo.z), o3 = unpack_bf16x2(uo.w);
      dsum += d0.x * o0.x + d0.y * o0.y + d1.x * o1.x + d1.y * o1.y + d2.x * o2.x + d2.y * o2.y + d3.x * o3.x +
              d3.y * o3.y;
    }
  }
  const float lse2 = row_ok ? lse[((long long)b * a.nh + h) * a.L + row] * kLog2e : 0.f;

  mbar_wait(&bars[1], 0);
  __syncwarp();
  tc_fence_after();

  // ---- P and dS for this thread's (row, column half)
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  const float sc2 = a.scale * kLog2e;
  const int groups = a.N16 >> 4;
  const int g_begin = half ? (groups + 1) / 2 : 0;
  const int g_end = half ? groups : (groups + 1) / 2;
  const int row8 = row & 7;
  const uint32_t prow_off = (row >> 3) * 1024 + row8 * 128;
  const unsigned long long drop_base =
      (((unsigned long long)b * a.nh + h) * a.L + row) * (unsigned long long)(a.P + a.L);
  for (int g = g_begin; g < g_end; ++g) {
    const int c = g << 4;
    uint32_t rs[16], rd[16];
    tmem_ld_32x32b_x16(t_row + c, rs);
    tmem_ld_32x32b_x16(t_row + DP_COL + c, rd);
    tmem_ld_wait();
    float p[16], ds[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float m = sMask[c + j];
      const bool ok = row_ok && (m != -INFINITY);
      float pj = exp2f(__uint_as_float(rs[j]) * sc2 + m * kLog2e - lse2);
      float dp = __uint_as_float(rd[j]);
      if (a.drop_thr) {
        const int ks = c + j;
        const int kk = ks < a.P8 ? ks : a.P + (ks - a.P8);          // reference key numbering
        const bool keep = dropout_keep(a.seed, drop_base + kk, a.drop_thr);
        dp = keep ? dp * a.drop_scale : 0.f;
        ds[j] = ok ? pj * (dp - dsum) * a.scale : 0.f;
        p[j] = (ok && keep) ? pj * a.drop_scale : 0.f;
      } else {
        ds[j] = ok ? pj * (dp - dsum) * a.scale : 0.f;
        p[j] = ok ? pj : 0.f;
      }
    }
#pragma unroll
    for (int gg = 0; gg < 2; ++gg) {
      const int key0 = c + gg * 8;
      const int chunk = key0 >> 6, c16 = (key0 & 63) >> 3;
      const uint32_t off = chunk * 16384 + prow_off + ((c16 ^ row8) << 4);
      uint4 u;
      u.x = pack_bf16x2(p[gg * 8 + 0], p[gg * 8 + 1]); u.y = pack_bf16x2(p[gg * 8 + 2], p[gg * 8 + 3]);
      u.z = pack_bf16x2(p[gg * 8 + 4], p[gg * 8 + 5]); u.w = pack_bf16x2(p[gg * 8 + 6], p[gg * 8 + 7]);
      *reinterpret_cast<uint4*>(sP + off) = u;
      u.x = pack_bf16x2(ds[gg * 8 + 0], ds[gg * 8 + 1]); u.y = pack_bf16x2(ds[gg * 8 + 2], ds[gg * 8 + 3]);
      u.z = pack_bf16x2(ds[gg * 8 + 4], ds[gg * 8 + 5]); u.w = pack_bf16x2(ds[gg * 8 + 6], ds[gg * 8 + 7]);
      *reinterpret_cast<uint4*>(sdS + off) = u;
    }
  }
  // key columns [N16, 64*n_chunks) of the last chunk are read by the key-tile MMAs as rows that are never
  // stored; they need no initialisation (TMEM lanes are independent).
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int n_tiles = (a.N16 + 127) / 128;               // 128-key tiles of dK / dV
  const int DQ_COL = 0, DK_COL = 64, DV_COL = DP_COL;     // dK tiles at 64, 128 ; dV tiles at 256, 320
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
    umma_commit(&bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);
  __syncwarp();
  tc_fence_after();

  // ---- stores: this thread owns 32 of the 64 head-dim columns of its row (TMEM loads are warp-collective:
  // all lanes issue them, only the global stores are predicated)
  const int dcol = half * 32;
  {
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_row + DQ_COL + dcol, r);
    tmem_ld_wait();
    if (row_ok) {
      __nv_bfloat16* o = dqkv + ((long long)b * a.L + row) * ld_dqkv + h * 64 + dcol;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
        u.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
        u.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
        u.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
        *reinterpret_cast<uint4*>(o + v * 8) = u;
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
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32b_x32(t_row + (which ? DV_COL : DK_COL) + t * 64 + dcol, r);
      tmem_ld_wait();
      if (is_text) {
        __nv_bfloat16* o = dqkv + ((long long)b * a.L + tx) * ld_dqkv + (which + 1) * H + h * 64 + dcol;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
          *reinterpret_cast<uint4*>(o + v * 8) = u;
        }
      } else if (is_prefix) {
        float* o = (which ? dvp : dkp);
        if (o) {
          o += (((long long)b * a.nh + h) * a.P + ks) * 64 + dcol;
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<float4*>(o + v * 4) =
                make_float4(__uint_as_float(r[v * 4 + 0]), __uint_as_float(r[v * 4 + 1]),
                            __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
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
  CUtensorMap tmdO, tmO;
  const uint64_t T = (uint64_t)a.B * a.L;
  const uint64_t H = (uint64_t)a.nh * 64;
  int rc = make_tmap_bf16_2d(&tmdO, dctx, H, T, ld_dctx, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmO, ctx, H, T, ld_ctx, 64, 128);
  if (rc) return rc;
  const AttnBwdSmem lay = attn_bwd_layout(a.P8, a.L64, a.N16);
  static bool set = false;
  if (!set) {
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  dim3 grid(1, a.nh, a.B);
  attn_bwd_tc_kernel<<<grid, kBwdThreads, lay.total, st>>>(m.q, m.kv, m.kp, m.vp, tmdO, tmO, a, lse,
                                                          (__nv_bfloat16*)dqkv, ld_dqkv, dkp, dvp);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtvaf
