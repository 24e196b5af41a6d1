// Prefix ("fusion") self-attention, fp32-math SIMT implementation (parity mode, any storage dtype).
//
// Restates RobertaSelfAttention.forward models/modeling_roberta.py:218-278 (BERT :282-333) without ever
// materialising the [B,nh,L,P+L] score tensor: the visual prefix K_p/V_p rows are streamed first, then
// the text K/V rows (the torch.cat of :221-222 becomes an index test), one online softmax over P+L
// keys, key-padding mask folded in as the additive -10000.0 of :1000, heads merged on store (:276-278).
// Backward = two kernels (dQ ; dK,dV incl. the gradient of the prefix) in the flash-attention form
// using the saved log-sum-exp.  Dropout on the probabilities (:268) is a counter-based hash keyed by
// (b,h,q,k) so backward regenerates the same mask.
#include "common.cuh"
#include "../../include/mtvaf_b200.h"
#include "attention_tc.cuh"

namespace mtvaf {

constexpr int AD = 64;          // head dim (fixed: 768/12 = 1024/16 = 64)
constexpr int AKT = 64;         // streamed rows per tile in fwd / dQ kernels
constexpr int APAD = 68;        // smem row stride (floats): 16-byte aligned, conflict-free for float4 rows
constexpr int AQB = 16;         // owner rows per block (4 warps x 4 rows)
constexpr float kMaskAdd = -10000.0f;

struct AttnArgs {
  const void* qkv; long long ld_qkv;
  const void* kp; const void* vp; int P;
  const long long* key_mask;
  int B, L, nh;
  float scale;
  uint32_t drop_thr; float drop_scale; unsigned long long seed;
  const unsigned long long* step;
};

template <typename T>
__device__ __forceinline__ void load_row64(const T* src, float* dst, int t4) {
  // 16 threads cooperate on a 64-float row: thread t4 in [0,16) moves 4 elements
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(dst + t4 * 4) = *reinterpret_cast<const float4*>(src + t4 * 4);
  } else {
    const uint2 u = *reinterpret_cast<const uint2*>(src + t4 * 4);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    *reinterpret_cast<float4*>(dst + t4 * 4) = make_float4(a.x, a.y, b.x, b.y);
  }
}

// pointer to the 64-wide K (which=1) or V (which=2) row of key `kk` (prefix rows first)
template <typename T>
__device__ __forceinline__ const T* kv_row(const AttnArgs& a, int b, int h, int kk, int which) {
  if (kk < a.P) {
    const T* base = reinterpret_cast<const T*>(which == 1 ? a.kp : a.vp);
    return base + (((long long)b * a.nh + h) * a.P + kk) * AD;
  }
  const T* base = reinterpret_cast<const T*>(a.qkv);
  return base + ((long long)b * a.L + (kk - a.P)) * a.ld_qkv + (long long)which * a.nh * AD + h * AD;
}

__device__ __forceinline__ float key_mask_add(const AttnArgs& a, int b, int kk) {
  if (kk < a.P) return 0.f;                                       // prefix mask is all ones (bert_model.py:491)
  return a.key_mask[(long long)b * a.L + (kk - a.P)] != 0 ? 0.f : kMaskAdd;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

// ================================================================ forward
template <typename T>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(AttnArgs a, T* __restrict__ ctx, long long ld_ctx, float* __restrict__ lse_out) {
  __shared__ __align__(16) float Qs[AQB][AD];
  __shared__ __align__(16) float Ks[AKT][APAD];
  __shared__ __align__(16) float Vs[AKT][APAD];
  __shared__ __align__(16) float Ps[4][4][AKT];
  __shared__ float Madd[AKT];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AQB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Lk = a.P + a.L;
  const T* qkv = reinterpret_cast<const T*>(a.qkv);

  for (int i = tid; i < AQB * 16; i += 128) {
    const int r = i >> 4, t4 = i & 15;
    const int q = q0 + r;
    if (q < a.L) {
      load_row64<T>(qkv + ((long long)b * a.L + q) * a.ld_qkv + h * AD, Qs[r], t4);
#pragma unroll
      for (int j = 0; j < 4; ++j) Qs[r][t4 * 4 + j] *= a.scale;
    } else {
      *reinterpret_cast<float4*>(&Qs[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float m[4], l[4], acc[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { m[j] = -INFINITY; l[j] = 0.f; acc[j][0] = 0.f; acc[j][1] = 0.f; }

  for (int k0 = 0; k0 < Lk; k0 += AKT) {
    __syncthreads();
    for (int i = tid; i < AKT * 16; i += 128) {
      const int r = i >> 4, t4 = i & 15;
      const int kk = k0 + r;
      if (kk < Lk) {
        load_row64<T>(kv_row<T>(a, b, h, kk, 1), Ks[r], t4);
        load_row64<T>(kv_row<T>(a, b, h, kk, 2), Vs[r], t4);
      } else {
        *reinterpret_cast<float4*>(&Ks[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(&Vs[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (tid < AKT) Madd[tid] = (k0 + tid < Lk) ? key_mask_add(a, b, k0 + tid) : -INFINITY;
    __syncthreads();

    float s[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { s[j][0] = 0.f; s[j][1] = 0.f; }
#pragma unroll 4
    for (int i = 0; i < AD; i += 4) {
      const float4 ka = *reinterpret_cast<const float4*>(&Ks[lane][i]);
      const float4 kb = *reinterpret_cast<const float4*>(&Ks[lane + 32][i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 qv = *reinterpret_cast<const float4*>(&Qs[warp * 4 + j][i]);
        s[j][0] = dot4(qv, ka, s[j][0]);
        s[j][1] = dot4(qv, kb, s[j][1]);
      }
    }
    const float ma = Madd[lane], mb = Madd[lane + 32];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s0 = s[j][0] + ma, s1 = s[j][1] + mb;
      const float mt = warp_max(fmaxf(s0, s1));
      const float mn = fmaxf(m[j], mt);
      const float corr = __expf(m[j] - mn);
      float p0 = __expf(s0 - mn), p1 = __expf(s1 - mn);
      l[j] = l[j] * corr + warp_sum(p0 + p1);
      acc[j][0] *= corr; acc[j][1] *= corr;
      m[j] = mn;
      if (a.drop_thr) {
        const int q = q0 + warp * 4 + j;
        const uint32_t rk = attn_drop_rowkey(step_seed(a.seed, a.step), ((unsigned long long)b * a.nh + h) * a.L + q);
        p0 = attn_drop_keep(rk, k0 + lane, a.drop_thr) ? p0 * a.drop_scale : 0.f;
        p1 = attn_drop_keep(rk, k0 + lane + 32, a.drop_thr) ? p1 * a.drop_scale : 0.f;
      }
      Ps[warp][j][lane] = p0;
      Ps[warp][j][lane + 32] = p1;
    }
    __syncwarp();
#pragma unroll 4
    for (int k = 0; k < AKT; k += 4) {
      float va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { va[u] = Vs[k + u][lane]; vb[u] = Vs[k + u][lane + 32]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 p = *reinterpret_cast<const float4*>(&Ps[warp][j][k]);
        acc[j][0] = fmaf(p.x, va[0], acc[j][0]); acc[j][0] = fmaf(p.y, va[1], acc[j][0]);
        acc[j][0] = fmaf(p.z, va[2], acc[j][0]); acc[j][0] = fmaf(p.w, va[3], acc[j][0]);
        acc[j][1] = fmaf(p.x, vb[0], acc[j][1]); acc[j][1] = fmaf(p.y, vb[1], acc[j][1]);
        acc[j][1] = fmaf(p.z, vb[2], acc[j][1]); acc[j][1] = fmaf(p.w, vb[3], acc[j][1]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = q0 + warp * 4 + j;
    if (q >= a.L) continue;
    const float inv = 1.f / l[j];
    T* o = ctx + ((long long)b * a.L + q) * ld_ctx + h * AD;
    o[lane] = from_f<T>(acc[j][0] * inv);
    o[lane + 32] = from_f<T>(acc[j][1] * inv);
    if (lane == 0) lse_out[((long long)b * a.nh + h) * a.L + q] = m[j] + __logf(l[j]);
  }
}

// attention probabilities (output_attentions=True, debug path): probs[b,h,q,k] = exp(s - lse)
template <typename T>
__global__ void attn_probs_kernel(AttnArgs a, const float* __restrict__ lse, float* __restrict__ probs) {
  const int Lk = a.P + a.L;
  const int b = blockIdx.z, h = blockIdx.y, q = blockIdx.x;
  __shared__ float qs[AD];
  const T* qkv = reinterpret_cast<const T*>(a.qkv);
  if (threadIdx.x < AD) qs[threadIdx.x] = to_f<T>(qkv[((long long)b * a.L + q) * a.ld_qkv + h * AD + threadIdx.x]) * a.scale;
  __syncthreads();
  const float ls = lse[((long long)b * a.nh + h) * a.L + q];
  for (int kk = threadIdx.x; kk < Lk; kk += blockDim.x) {
    const T* kr = kv_row<T>(a, b, h, kk, 1);
    float s = 0.f;
    for (int i = 0; i < AD; ++i) s = fmaf(qs[i], to_f<T>(kr[i]), s);
    s += key_mask_add(a, b, kk);
    probs[((((long long)b * a.nh + h) * a.L + q) * Lk) + kk] = __expf(s - ls);
  }
}

// ================================================================ backward
// dsum[b,h,q] = sum_d dO[q,d] * O[q,d]
template <typename T>
__global__ void attn_dsum_kernel(const T* __restrict__ dctx, long long ld_d, const T* __restrict__ ctx, long long ld_c,
                                 int B, int L, int nh, float* __restrict__ dsum) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = B * nh * L;
  if (w >= total) return;
  const int q = w % L, h = (w / L) % nh, b = w / (L * nh);
  const T* d = dctx + ((long long)b * L + q) * ld_d + h * AD;
  const T* o = ctx + ((long long)b * L + q) * ld_c + h * AD;
  float s = to_f<T>(d[lane]) * to_f<T>(o[lane]) + to_f<T>(d[lane + 32]) * to_f<T>(o[lane + 32]);
  s = warp_sum(s);
  if (lane == 0) dsum[w] = s;
}

// dQ: block owns 16 queries, streams key tiles
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(AttnArgs a, const T* __restrict__ dctx, long long ld_d, const float* __restrict__ lse,
                   const float* __restrict__ dsum, T* __restrict__ dqkv, long long ld_dqkv) {
  __shared__ __align__(16) float Qs[AQB][AD];
  __shared__ __align__(16) float Ds[AQB][AD];
  __shared__ __align__(16) float Ks[AKT][APAD];
  __shared__ __align__(16) float Vs[AKT][APAD];
  __shared__ __align__(16) float Ps[4][4][AKT];
  __shared__ float Madd[AKT];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AQB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Lk = a.P + a.L;
  const T* qkv = reinterpret_cast<const T*>(a.qkv);
  for (int i = tid; i < AQB * 16; i += 128) {
    const int r = i >> 4, t4 = i & 15;
    const int q = q0 + r;
    if (q < a.L) {
      load_row64<T>(qkv + ((long long)b * a.L + q) * a.ld_qkv + h * AD, Qs[r], t4);
      load_row64<T>(dctx + ((long long)b * a.L + q) * ld_d + h * AD, Ds[r], t4);
#pragma unroll
      for (int j = 0; j < 4; ++j) Qs[r][t4 * 4 + j] *= a.scale;
    } else {
      *reinterpret_cast<float4*>(&Qs[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&Ds[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float ls[4], dsm[4], acc[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = q0 + warp * 4 + j;
    const long long idx = ((long long)b * a.nh + h) * a.L + q;
    ls[j] = (q < a.L) ? lse[idx] : INFINITY;
    dsm[j] = (q < a.L) ? dsum[idx] : 0.f;
    acc[j][0] = 0.f; acc[j][1] = 0.f;
  }
  for (int k0 = 0; k0 < Lk; k0 += AKT) {
    __syncthreads();
    for (int i = tid; i < AKT * 16; i += 128) {
      const int r = i >> 4, t4 = i & 15;
      const int kk = k0 + r;
      if (kk < Lk) {
        load_row64<T>(kv_row<T>(a, b, h, kk, 1), Ks[r], t4);
        load_row64<T>(kv_row<T>(a, b, h, kk, 2), Vs[r], t4);
      } else {
        *reinterpret_cast<float4*>(&Ks[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(&Vs[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (tid < AKT) Madd[tid] = (k0 + tid < Lk) ? key_mask_add(a, b, k0 + tid) : -INFINITY;
    __syncthreads();
    float s[4][2], dp[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { s[j][0] = 0.f; s[j][1] = 0.f; dp[j][0] = 0.f; dp[j][1] = 0.f; }
#pragma unroll 2
    for (int i = 0; i < AD; i += 4) {
      const float4 ka = *reinterpret_cast<const float4*>(&Ks[lane][i]);
      const float4 kb = *reinterpret_cast<const float4*>(&Ks[lane + 32][i]);
      const float4 va = *reinterpret_cast<const float4*>(&Vs[lane][i]);
      const float4 vb = *reinterpret_cast<const float4*>(&Vs[lane + 32][i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 qv = *reinterpret_cast<const float4*>(&Qs[warp * 4 + j][i]);
        const float4 dv = *reinterpret_cast<const float4*>(&Ds[warp * 4 + j][i]);
        s[j][0] = dot4(qv, ka, s[j][0]);
        s[j][1] = dot4(qv, kb, s[j][1]);
        dp[j][0] = dot4(dv, va, dp[j][0]);
        dp[j][1] = dot4(dv, vb, dp[j][1]);
      }
    }
    const float ma = Madd[lane], mb = Madd[lane + 32];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p0 = __expf(s[j][0] + ma - ls[j]), p1 = __expf(s[j][1] + mb - ls[j]);
      float d0 = dp[j][0], d1 = dp[j][1];
      if (a.drop_thr) {
        const int q = q0 + warp * 4 + j;
        const uint32_t rk = attn_drop_rowkey(step_seed(a.seed, a.step), ((unsigned long long)b * a.nh + h) * a.L + q);
        d0 = attn_drop_keep(rk, k0 + lane, a.drop_thr) ? d0 * a.drop_scale : 0.f;
        d1 = attn_drop_keep(rk, k0 + lane + 32, a.drop_thr) ? d1 * a.drop_scale : 0.f;
      }
      Ps[warp][j][lane] = p0 * (d0 - dsm[j]);
      Ps[warp][j][lane + 32] = p1 * (d1 - dsm[j]);
    }
    __syncwarp();
#pragma unroll 4
    for (int k = 0; k < AKT; k += 4) {
      float ka[4], kb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { ka[u] = Ks[k + u][lane]; kb[u] = Ks[k + u][lane + 32]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 p = *reinterpret_cast<const float4*>(&Ps[warp][j][k]);
        acc[j][0] = fmaf(p.x, ka[0], acc[j][0]); acc[j][0] = fmaf(p.y, ka[1], acc[j][0]);
        acc[j][0] = fmaf(p.z, ka[2], acc[j][0]); acc[j][0] = fmaf(p.w, ka[3], acc[j][0]);
        acc[j][1] = fmaf(p.x, kb[0], acc[j][1]); acc[j][1] = fmaf(p.y, kb[1], acc[j][1]);
        acc[j][1] = fmaf(p.z, kb[2], acc[j][1]); acc[j][1] = fmaf(p.w, kb[3], acc[j][1]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int q = q0 + warp * 4 + j;
    if (q >= a.L) continue;
    T* o = dqkv + ((long long)b * a.L + q) * ld_dqkv + h * AD;
    o[lane] = from_f<T>(acc[j][0] * a.scale);
    o[lane + 32] = from_f<T>(acc[j][1] * a.scale);
  }
}

// dK, dV: block owns 16 keys (prefix rows and text rows alike), streams 32-query tiles
template <typename T>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(AttnArgs a, const T* __restrict__ dctx, long long ld_d, const float* __restrict__ lse,
                    const float* __restrict__ dsum, T* __restrict__ dqkv, long long ld_dqkv, float* __restrict__ dkp,
                    float* __restrict__ dvp) {
  constexpr int QT = 32;
  __shared__ __align__(16) float Ko[AQB][AD];
  __shared__ __align__(16) float Vo[AQB][AD];
  __shared__ __align__(16) float Qs[QT][APAD];
  __shared__ __align__(16) float Ds[QT][APAD];
  __shared__ __align__(16) float Pd[4][4][QT];    // ds
  __shared__ __align__(16) float Pp[4][4][QT];    // dropped probabilities
  __shared__ float Ls[QT], Dm[QT];
  const int b = blockIdx.z, h = blockIdx.y, kbase = blockIdx.x * AQB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Lk = a.P + a.L;
  const T* qkv = reinterpret_cast<const T*>(a.qkv);
  for (int i = tid; i < AQB * 16; i += 128) {
    const int r = i >> 4, t4 = i & 15;
    const int kk = kbase + r;
    if (kk < Lk) {
      load_row64<T>(kv_row<T>(a, b, h, kk, 1), Ko[r], t4);
      load_row64<T>(kv_row<T>(a, b, h, kk, 2), Vo[r], t4);
    } else {
      *reinterpret_cast<float4*>(&Ko[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(&Vo[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float madd[4], dk[4][2], dv[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kk = kbase + warp * 4 + j;
    madd[j] = (kk < Lk) ? key_mask_add(a, b, kk) : -INFINITY;
    dk[j][0] = dk[j][1] = dv[j][0] = dv[j][1] = 0.f;
  }
  for (int q0 = 0; q0 < a.L; q0 += QT) {
    __syncthreads();
    for (int i = tid; i < QT * 16; i += 128) {
      const int r = i >> 4, t4 = i & 15;
      const int q = q0 + r;
      if (q < a.L) {
        load_row64<T>(qkv + ((long long)b * a.L + q) * a.ld_qkv + h * AD, Qs[r], t4);
        load_row64<T>(dctx + ((long long)b * a.L + q) * ld_d + h * AD, Ds[r], t4);
#pragma unroll
        for (int j = 0; j < 4; ++j) Qs[r][t4 * 4 + j] *= a.scale;
      } else {
        *reinterpret_cast<float4*>(&Qs[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(&Ds[r][t4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (tid < QT) {
      const int q = q0 + tid;
      const long long idx = ((long long)b * a.nh + h) * a.L + q;
      Ls[tid] = (q < a.L) ? lse[idx] : INFINITY;
      Dm[tid] = (q < a.L) ? dsum[idx] : 0.f;
    }
    __syncthreads();
    float s[4], dp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { s[j] = 0.f; dp[j] = 0.f; }
#pragma unroll 4
    for (int i = 0; i < AD; i += 4) {
      const float4 qv = *reinterpret_cast<const float4*>(&Qs[lane][i]);
      const float4 dd = *reinterpret_cast<const float4*>(&Ds[lane][i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 kv = *reinterpret_cast<const float4*>(&Ko[warp * 4 + j][i]);
        const float4 vv = *reinterpret_cast<const float4*>(&Vo[warp * 4 + j][i]);
        s[j] = dot4(qv, kv, s[j]);
        dp[j] = dot4(dd, vv, dp[j]);
      }
    }
    const float lsq = Ls[lane], dmq = Dm[lane];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p = __expf(s[j] + madd[j] - lsq);
      float d = dp[j], pd = p;
      if (a.drop_thr) {
        const int q = q0 + lane, kk = kbase + warp * 4 + j;
        const bool keep = attn_drop_keep(attn_drop_rowkey(step_seed(a.seed, a.step), ((unsigned long long)b * a.nh + h) * a.L + q), kk,
                                         a.drop_thr);
        d = keep ? d * a.drop_scale : 0.f;
        pd = keep ? p * a.drop_scale : 0.f;
      }
      Pd[warp][j][lane] = p * (d - dmq);
      Pp[warp][j][lane] = pd;
    }
    __syncwarp();
#pragma unroll 4
    for (int q = 0; q < QT; q += 4) {
      float qa[4], qb[4], da[4], db[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        qa[u] = Qs[q + u][lane]; qb[u] = Qs[q + u][lane + 32];
        da[u] = Ds[q + u][lane]; db[u] = Ds[q + u][lane + 32];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 ds = *reinterpret_cast<const float4*>(&Pd[warp][j][q]);
        const float4 pp = *reinterpret_cast<const float4*>(&Pp[warp][j][q]);
        dk[j][0] = fmaf(ds.x, qa[0], dk[j][0]); dk[j][0] = fmaf(ds.y, qa[1], dk[j][0]);
        dk[j][0] = fmaf(ds.z, qa[2], dk[j][0]); dk[j][0] = fmaf(ds.w, qa[3], dk[j][0]);
        dk[j][1] = fmaf(ds.x, qb[0], dk[j][1]); dk[j][1] = fmaf(ds.y, qb[1], dk[j][1]);
        dk[j][1] = fmaf(ds.z, qb[2], dk[j][1]); dk[j][1] = fmaf(ds.w, qb[3], dk[j][1]);
        dv[j][0] = fmaf(pp.x, da[0], dv[j][0]); dv[j][0] = fmaf(pp.y, da[1], dv[j][0]);
        dv[j][0] = fmaf(pp.z, da[2], dv[j][0]); dv[j][0] = fmaf(pp.w, da[3], dv[j][0]);
        dv[j][1] = fmaf(pp.x, db[0], dv[j][1]); dv[j][1] = fmaf(pp.y, db[1], dv[j][1]);
        dv[j][1] = fmaf(pp.z, db[2], dv[j][1]); dv[j][1] = fmaf(pp.w, db[3], dv[j][1]);
      }
    }
    __syncwarp();
  }
  // Qs already carries the 1/sqrt(d) scale, so dk needs no further scaling
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kk = kbase + warp * 4 + j;
    if (kk >= Lk) continue;
    if (kk < a.P) {
      if (dkp) {
        float* o = dkp + (((long long)b * a.nh + h) * a.P + kk) * AD;
        o[lane] = dk[j][0]; o[lane + 32] = dk[j][1];
      }
      if (dvp) {
        float* o = dvp + (((long long)b * a.nh + h) * a.P + kk) * AD;
        o[lane] = dv[j][0]; o[lane + 32] = dv[j][1];
      }
    } else {
      T* o = dqkv + ((long long)b * a.L + (kk - a.P)) * ld_dqkv + h * AD;
      const long long H = (long long)a.nh * AD;
      o[H + lane] = from_f<T>(dk[j][0]); o[H + lane + 32] = from_f<T>(dk[j][1]);
      o[2 * H + lane] = from_f<T>(dv[j][0]); o[2 * H + lane + 32] = from_f<T>(dv[j][1]);
    }
  }
}

static int fill_args(AttnArgs* a, const void* qkv, int64_t ld_qkv, const void* kp, const void* vp, int P,
                     const int64_t* key_mask, int B, int L, int nh, int d, float p_drop, uint64_t seed, int dtype) {
  MTVAF_REQUIRE(qkv && key_mask, "attention: null argument");
  MTVAF_REQUIRE(d == AD, "attention: head dim %d unsupported (only 64)", d);
  MTVAF_REQUIRE(P == 0 || (kp && vp), "attention: prefix pointers missing");
  MTVAF_REQUIRE(B > 0 && L > 0 && nh > 0 && P >= 0, "attention: bad shape");
  const int al = (dtype == MTVAF_BF16) ? 4 : 4;   // 4-element vector loads: 8 B (bf16) / 16 B (fp32)
  MTVAF_REQUIRE(ld_qkv % al == 0, "attention: ld_qkv must be a multiple of 4");
  a->qkv = qkv; a->ld_qkv = ld_qkv; a->kp = kp; a->vp = vp; a->P = P;
  a->key_mask = reinterpret_cast<const long long*>(key_mask);
  a->B = B; a->L = L; a->nh = nh;
  a->scale = 1.0f / sqrtf((float)d);
  a->drop_thr = 0; a->drop_scale = 1.f; a->seed = seed; a->step = step_source();
  if (p_drop > 0.f) {
    MTVAF_REQUIRE(p_drop < 1.f, "attention: dropout p must be < 1");
    double t = (double)p_drop * 4294967296.0;
    a->drop_thr = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    a->drop_scale = 1.f / (1.f - p_drop);
  }
  return 0;
}

}  // namespace mtvaf

using namespace mtvaf;

static int g_attention_impl = 0;
namespace mtvaf { int attention_impl_override() { return g_attention_impl; } }
extern "C" int mtvaf_set_attention_impl(int impl) {
  MTVAF_REQUIRE(impl >= 0 && impl <= 2,
                "attention impl must be 0 (auto), 1 (SIMT) or 2 (tcgen05, generic instead of the pipelined backward)");
  g_attention_impl = impl;
  return 0;
}

// bytes of caller-provided workspace mtvaf_attention_fwd_ws needs for this shape (0 for most: only long bf16 text whose
// keys do not fit one resident-key tile set -- P + L > ~400 -- runs as two key windows + a merge)
extern "C" int64_t mtvaf_attention_fwd_workspace_bytes(int B, int L, int nh, int d, int P, int dtype) {
  if (dtype != MTVAF_BF16 || d != 64 || L <= 256 || L > 512) return 0;
  AttnTcArgs ta;
  ta.P = P; ta.P8 = (P + 7) / 8 * 8; ta.L = L; ta.L64 = (L + 63) / 64 * 64; ta.N16 = (ta.P8 + L + 15) / 16 * 16;
  ta.Lk = L; ta.kt0 = 0; ta.kbase = P; ta.B = B; ta.nh = nh;
  if (attn_fwd_tc_fits(ta) || !attn_fwd_tc_windows_supported(ta)) return 0;
  return (int64_t)attn_fwd_tc_windows_workspace(B, L, nh);
}

extern "C" int mtvaf_attention_fwd(const void* qkv, int64_t ld_qkv, const void* kp, const void* vp, int P,
                                   const int64_t* key_mask, int B, int L, int nh, int d, void* ctx, int64_t ld_ctx,
                                   float* lse, float* probs, int dtype, float p_drop, uint64_t seed, void* stream) {
  return mtvaf_attention_fwd_ws(qkv, ld_qkv, kp, vp, P, key_mask, B, L, nh, d, ctx, ld_ctx, lse, probs, dtype, p_drop,
                                seed, nullptr, 0, stream);
}

extern "C" int mtvaf_attention_fwd_ws(const void* qkv, int64_t ld_qkv, const void* kp, const void* vp, int P,
                                      const int64_t* key_mask, int B, int L, int nh, int d, void* ctx, int64_t ld_ctx,
                                      float* lse, float* probs, int dtype, float p_drop, uint64_t seed,
                                      void* workspace, int64_t workspace_bytes, void* stream) {
  AttnArgs a;
  if (int rc = fill_args(&a, qkv, ld_qkv, kp, vp, P, key_mask, B, L, nh, d, p_drop, seed, dtype)) return rc;
  MTVAF_REQUIRE(ctx && lse, "attention_fwd: null output");
  dim3 grid((L + AQB - 1) / AQB, nh, B);
  cudaStream_t st = (cudaStream_t)stream;
  bool done = false;
  if (dtype == MTVAF_BF16 && attention_impl_override() != 1 && ld_ctx % 8 == 0) {
    // tensor-core path (tcgen05): every shape of the benchmark configs; others fall through to SIMT
    AttnTcArgs ta;
    AttnTcMaps tm;
    bool ok = false;
    if (int rc = attn_tc_prepare(qkv, ld_qkv, kp, vp, P, key_mask, B, L, nh, p_drop, seed, &ta, &tm, &ok)) return rc;
    if (ok && attn_fwd_tc_fits(ta)) {
      if (int rc = attn_fwd_tc_launch(ta, tm, ctx, ld_ctx, lse, st)) return rc;
      done = true;
    } else if (ok && attn_fwd_tc_windows_supported(ta) && workspace &&
               workspace_bytes >= (int64_t)attn_fwd_tc_windows_workspace(B, L, nh) &&
               (reinterpret_cast<uintptr_t>(workspace) & 255) == 0) {
      // long text: the keys of an item do not fit -> two key windows through the same kernel, then a merge
      if (int rc = attn_fwd_tc_windows_launch(ta, tm, ctx, ld_ctx, lse, workspace, st)) return rc;
      done = true;
    }
  }
  if (!done) {
    if (dtype == MTVAF_BF16) attn_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(a, (__nv_bfloat16*)ctx, ld_ctx, lse);
    else attn_fwd_kernel<float><<<grid, 128, 0, st>>>(a, (float*)ctx, ld_ctx, lse);
    MTVAF_LAUNCH_CHECK();
  }
  if (probs) {
    dim3 g2(L, nh, B);
    if (dtype == MTVAF_BF16) attn_probs_kernel<__nv_bfloat16><<<g2, 128, 0, st>>>(a, lse, probs);
    else attn_probs_kernel<float><<<g2, 128, 0, st>>>(a, lse, probs);
    MTVAF_LAUNCH_CHECK();
  }
  return 0;
}

// d_bias_qkv (optional): fp32 [3 * nh * d], += column sums of dqkv -- the bias gradient of the fused QKV projection
// (models/modeling_roberta.py:202,219-220).  The pipelined tcgen05 kernel produces it while draining TMEM; every other
// path runs the column-sum kernel over dqkv afterwards, so callers see one behaviour.
extern "C" int mtvaf_attention_bwd_ex(const void* dctx, int64_t ld_dctx, const void* qkv, int64_t ld_qkv,
                                      const void* kp, const void* vp, int P, const int64_t* key_mask, const void* ctx,
                                      int64_t ld_ctx, const float* lse, int B, int L, int nh, int d, void* dqkv,
                                      int64_t ld_dqkv, float* dkp, float* dvp, float* dsum_scratch, int dtype,
                                      float p_drop, uint64_t seed, float* d_bias_qkv, void* stream) {
  AttnArgs a;
  if (int rc = fill_args(&a, qkv, ld_qkv, kp, vp, P, key_mask, B, L, nh, d, p_drop, seed, dtype)) return rc;
  MTVAF_REQUIRE(dctx && ctx && lse && dqkv && dsum_scratch, "attention_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int total = B * nh * L;
  dim3 gq((L + AQB - 1) / AQB, nh, B), gk((P + L + AQB - 1) / AQB, nh, B);
  if (dtype == MTVAF_BF16 && attention_impl_override() != 1 && ld_ctx % 8 == 0 && ld_dctx % 8 == 0 &&
      ld_dqkv % 8 == 0 && (reinterpret_cast<uintptr_t>(dctx) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0) {
    // tensor-core path (tcgen05): one CTA per (batch, head), L <= 128
    AttnTcArgs ta;
    AttnTcMaps tm;
    bool ok = false;
    if (int rc = attn_tc_prepare(qkv, ld_qkv, kp, vp, P, key_mask, B, L, nh, p_drop, seed, &ta, &tm, &ok)) return rc;
    if (ok && attention_impl_override() == 0 && attn_bwd_pair_supported(ta))
      return attn_bwd_pair_launch(ta, tm, dctx, ld_dctx, ctx, ld_ctx, lse, dqkv, ld_dqkv, dkp, dvp, d_bias_qkv, st);
    if (ok && attention_impl_override() == 0 && attn_bwd_pipe_supported(ta))
      return attn_bwd_pipe_launch(ta, tm, dctx, ld_dctx, ctx, ld_ctx, lse, dqkv, ld_dqkv, dkp, dvp, d_bias_qkv, st);
    if (ok && attn_bwd_tc_supported(ta)) {
      if (int rc = attn_bwd_tc_launch(ta, tm, dctx, ld_dctx, ctx, ld_ctx, lse, dqkv, ld_dqkv, dkp, dvp, st)) return rc;
      return d_bias_qkv ? mtvaf_colsum(dqkv, ld_dqkv, dtype, B * L, 3 * nh * d, d_bias_qkv, stream) : 0;
    }
    if (ok && attn_bwd_long_supported(ta)) {                                        // 128 < L <= 512
      if (int rc = attn_bwd_long_launch(ta, tm, dctx, ld_dctx, ctx, ld_ctx, lse, dqkv, ld_dqkv, dkp, dvp, st)) return rc;
      return d_bias_qkv ? mtvaf_colsum(dqkv, ld_dqkv, dtype, B * L, 3 * nh * d, d_bias_qkv, stream) : 0;
    }
  }
  if (dtype == MTVAF_BF16) {
    using T = __nv_bfloat16;
    attn_dsum_kernel<T><<<(total * 32 + 255) / 256, 256, 0, st>>>((const T*)dctx, ld_dctx, (const T*)ctx, ld_ctx, B, L, nh, dsum_scratch);
    MTVAF_LAUNCH_CHECK();
    attn_bwd_dq_kernel<T><<<gq, 128, 0, st>>>(a, (const T*)dctx, ld_dctx, lse, dsum_scratch, (T*)dqkv, ld_dqkv);
    MTVAF_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<T><<<gk, 128, 0, st>>>(a, (const T*)dctx, ld_dctx, lse, dsum_scratch, (T*)dqkv, ld_dqkv, dkp, dvp);
  } else {
    using T = float;
    attn_dsum_kernel<T><<<(total * 32 + 255) / 256, 256, 0, st>>>((const T*)dctx, ld_dctx, (const T*)ctx, ld_ctx, B, L, nh, dsum_scratch);
    MTVAF_LAUNCH_CHECK();
    attn_bwd_dq_kernel<T><<<gq, 128, 0, st>>>(a, (const T*)dctx, ld_dctx, lse, dsum_scratch, (T*)dqkv, ld_dqkv);
    MTVAF_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<T><<<gk, 128, 0, st>>>(a, (const T*)dctx, ld_dctx, lse, dsum_scratch, (T*)dqkv, ld_dqkv, dkp, dvp);
  }
  MTVAF_LAUNCH_CHECK();
  return d_bias_qkv ? mtvaf_colsum(dqkv, ld_dqkv, dtype, B * L, 3 * nh * d, d_bias_qkv, stream) : 0;
}

extern "C" int mtvaf_attention_bwd(const void* dctx, int64_t ld_dctx, const void* qkv, int64_t ld_qkv,
                                   const void* kp, const void* vp, int P, const int64_t* key_mask, const void* ctx,
                                   int64_t ld_ctx, const float* lse, int B, int L, int nh, int d, void* dqkv,
                                   int64_t ld_dqkv, float* dkp, float* dvp, float* dsum_scratch, int dtype,
                                   float p_drop, uint64_t seed, void* stream) {
  return mtvaf_attention_bwd_ex(dctx, ld_dctx, qkv, ld_qkv, kp, vp, P, key_mask, ctx, ld_ctx, lse, B, L, nh, d, dqkv,
                                ld_dqkv, dkp, dvp, dsum_scratch, dtype, p_drop, seed, nullptr, stream);
}
