// Prefix ("fusion") self-attention on the tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
//
// One CTA (128 threads, thread == query row == TMEM lane) per (batch, head, 128-query tile):
//   TMA  : Q tile, the visual-prefix K_p/V_p rows (8-row boxes) and the text K/V rows (64-row boxes)
//          land in SWIZZLE_128B shared memory -- the torch.cat of models/modeling_roberta.py:221-222 is
//          just two groups of TMA boxes into one key-row numbering (prefix padded to a multiple of 8).
//   MMA 1: S[q, key] = Q K^T  (tcgen05.mma, both operands K-major)            -> TMEM columns [0, N16)
//   SIMT : each thread reads its row of S from TMEM, applies 1/sqrt(d) and the additive key mask
//          (-10000.0, models/modeling_roberta.py:1000), softmax (two passes over TMEM), optional dropout,
//          and writes P as bf16 into shared memory in the K-major SWIZZLE_128B operand layout
//   MMA 2: O[q, d] = P V      (A = P K-major from smem, B = V MN-major)       -> TMEM columns [S_COLS, +64)
//   store: O / rowsum -> ctx (heads merged, :276-278), lse = max + log(rowsum) saved for backward.
// Small smem/TMEM footprint lets 2 CTAs share an SM so one CTA's softmax overlaps the other's MMAs.
// The backward kernel (same tiling, five MMAs) is in attention_tc_bwd.cu.
#include "attention_tc.cuh"

namespace mtvaf {
using namespace ptx;

template <bool BIG>
__global__ void __launch_bounds__(128, BIG ? 1 : 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmKp, const __grid_constant__ CUtensorMap tmVp,
                   AttnTcArgs a, __nv_bfloat16* __restrict__ ctx, long long ld_ctx, float* __restrict__ lse_out) {
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
  float* sMask = reinterpret_cast<float*>(sP + n_chunks * 16384);     // [N16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + ((a.N16 + 15) / 16) * 16);   // load, s, o
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int H = a.nh * 64;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmKV);
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  __syncwarp();                                      // tcgen05.alloc is .sync.aligned: reconverge warp 0 first
  if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_ptr);
  // additive key mask in smem-key numbering: prefix rows [0,P) visible, [P,P8) padding, text rows follow
  for (int k = tid; k < a.N16; k += 128) {
    float m;
    if (k < a.P8) m = (k < a.P) ? 0.f : -INFINITY;
    else {
      const int t = k - a.P8;
      m = (t < a.L) ? (a.key_mask[(long long)b * a.L + t] != 0 ? 0.f : -10000.0f) : -INFINITY;
    }
    sMask[k] = m;
  }
  // K rows [key_rows, N16) alias the first V rows and V rows [key_rows, N16) alias the first rows of sP:
  // both hold finite bf16 values, and those key columns get p = 0 (mask -inf), so they contribute nothing.
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (tid == 0) {
    const uint32_t bytes = 16384u + 2u * key_rows * 128u;
    mbar_arrive_expect_tx(&bars[0], bytes);
    tma_load_2d(sQ, &tmQ, &bars[0], h * 64, b * a.L + q0);
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
  __syncwarp();
  mbar_wait(&bars[1], 0);
  __syncwarp();
  tc_fence_after();

  // ---- softmax over this thread's row of S (two passes over TMEM)
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const float sc = a.scale * 1.4426950408889634f;          // fold log2(e): exp(x) = exp2(x * log2e)
  float mx = -INFINITY;
  for (int c = 0; c < a.N16; c += 16) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(t_row + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(r[j]) * a.scale + sMask[c + j]);
  }
  const int q = q0 + tid;
  const int row8 = tid & 7;
  uint8_t* prow = sP + (tid >> 3) * 1024 + row8 * 128;
  float sum = 0.f;
  const float mx2 = mx * 1.4426950408889634f;
  const unsigned long long drop_base =
      (((unsigned long long)b * a.nh + h) * a.L + q) * (unsigned long long)(a.P + a.L);
  for (int c = 0; c < a.N16; c += 16) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(t_row + c, r);
    tmem_ld_wait();
    float p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x = __uint_as_float(r[j]) * sc + sMask[c + j] * 1.4426950408889634f - mx2;
      p[j] = exp2f(x);
      sum += p[j];
    }
    if (a.drop_thr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        // dropout index uses the reference key numbering (prefix 0..P-1, text P..P+L-1)
        const int ks = c + j;
        const int kk = ks < a.P8 ? ks : a.P + (ks - a.P8);
        p[j] = dropout_keep(a.seed, drop_base + kk, a.drop_thr) ? p[j] * a.drop_scale : 0.f;
      }
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int key0 = c + g * 8;
      const int chunk = key0 >> 6, c16 = (key0 & 63) >> 3;
      uint4 u;
      u.x = pack_bf16x2(p[g * 8 + 0], p[g * 8 + 1]); u.y = pack_bf16x2(p[g * 8 + 2], p[g * 8 + 3]);
      u.z = pack_bf16x2(p[g * 8 + 4], p[g * 8 + 5]); u.w = pack_bf16x2(p[g * 8 + 6], p[g * 8 + 7]);
      *reinterpret_cast<uint4*>(prow + chunk * 16384 + ((c16 ^ row8) << 4)) = u;
    }
  }
  fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    // ---- O = P V
    const uint32_t aP = smem_u32(sP), aV = smem_u32(sV);
    const uint32_t idesc = make_idesc_bf16(128, 64, false, true);
    const int ksteps = a.N16 / 16;
    for (int j = 0; j < ksteps; ++j)
      umma_f16_ss(tmem_base + S_COLS, make_smem_desc_sw128(aP + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                  make_smem_desc_sw128(aV + j * 2048, 8192, 1024), idesc, j > 0 ? 1u : 0u);
    umma_commit(&bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);
  __syncwarp();
  tc_fence_after();

  // tcgen05.ld is warp-collective (.sync.aligned): EVERY lane issues the loads (rows past L included,
  // never inside a divergent branch); only the global stores are predicated.
  const float inv = 1.f / sum;
  const bool valid = q < a.L;
  __nv_bfloat16* o = ctx + ((long long)b * a.L + (valid ? q : 0)) * ld_ctx + h * 64;
#pragma unroll
  for (int c = 0; c < 64; c += 16) {
    uint32_t r[16];
    __syncwarp();
    tmem_ld_32x32b_x16(t_row + S_COLS + c, r);
    tmem_ld_wait();
    uint4 u0, u1;
    u0.x = pack_bf16x2(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
    u0.y = pack_bf16x2(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
    u0.z = pack_bf16x2(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
    u0.w = pack_bf16x2(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
    u1.x = pack_bf16x2(__uint_as_float(r[8]) * inv, __uint_as_float(r[9]) * inv);
    u1.y = pack_bf16x2(__uint_as_float(r[10]) * inv, __uint_as_float(r[11]) * inv);
    u1.z = pack_bf16x2(__uint_as_float(r[12]) * inv, __uint_as_float(r[13]) * inv);
    u1.w = pack_bf16x2(__uint_as_float(r[14]) * inv, __uint_as_float(r[15]) * inv);
    if (valid) {
      *reinterpret_cast<uint4*>(o + c) = u0;
      *reinterpret_cast<uint4*>(o + c + 8) = u1;
    }
  }
  if (valid) lse_out[((long long)b * a.nh + h) * a.L + q] = mx + __logf(sum);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

size_t attn_fwd_tc_smem(const AttnTcArgs& a) {
  const int key_rows = a.P8 + a.L64;
  const int n_chunks = (a.N16 + 63) / 64;
  return 1024 + 16384 + 2 * (size_t)key_rows * 128 + (size_t)n_chunks * 16384 + ((a.N16 + 15) / 16) * 64 + 64;
}

int attn_fwd_tc_launch(const AttnTcArgs& a, const AttnTcMaps& m, void* ctx, int64_t ld_ctx, float* lse,
                       cudaStream_t st) {
  const size_t smem = attn_fwd_tc_smem(a);
  dim3 grid((a.L + 127) / 128, a.nh, a.B);
  const bool big = a.N16 > 192;
  static bool set0 = false, set1 = false;
  if (!big) {
    if (!set0) {
      MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set0 = true;
    }
    attn_fwd_tc_kernel<false><<<grid, 128, smem, st>>>(m.q, m.kv, m.kp, m.vp, a, (__nv_bfloat16*)ctx, ld_ctx, lse);
  } else {
    if (!set1) {
      MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      set1 = true;
    }
    attn_fwd_tc_kernel<true><<<grid, 128, smem, st>>>(m.q, m.kv, m.kp, m.vp, a, (__nv_bfloat16*)ctx, ld_ctx, lse);
  }
  MTVAF_LAUNCH_CHECK();
  return 0;
}

// shape gate + tensor maps shared by forward and backward
int attn_tc_prepare(const void* qkv, int64_t ld_qkv, const void* kp, const void* vp, int P, const int64_t* key_mask,
                    int B, int L, int nh, float p_drop, uint64_t seed, AttnTcArgs* a, AttnTcMaps* m, bool* ok) {
  *ok = false;
  a->P = P; a->P8 = (P + 7) / 8 * 8; a->L = L; a->L64 = (L + 63) / 64 * 64;
  a->N16 = (a->P8 + L + 15) / 16 * 16;
  a->B = B; a->nh = nh; a->key_mask = reinterpret_cast<const long long*>(key_mask);
  a->scale = 0.125f;
  a->drop_thr = 0; a->drop_scale = 1.f; a->seed = seed;
  if (p_drop > 0.f) {
    double t = (double)p_drop * 4294967296.0;
    a->drop_thr = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    a->drop_scale = 1.f / (1.f - p_drop);
  }
  if (a->N16 > 448 || ld_qkv % 8 != 0 || (reinterpret_cast<uintptr_t>(qkv) & 15)) return 0;
  if (attn_fwd_tc_smem(*a) > 227 * 1024) return 0;
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
