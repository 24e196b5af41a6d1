// LayerNorm forward/backward and the embedding gather + LayerNorm (+ bit-exact RoBERTa position ids).
// HBM-bound: one warp per row, 16-byte vector loads held in registers, warp-shuffle reductions
// (mean first, then centred variance -- the same two-pass form torch uses), fp32 statistics saved
// for backward.  Replaces nn.LayerNorm / nn.Embedding call sites models/modeling_roberta.py:131-139,
// 298,381 and models/modeling_bert.py:212-221.
#include "common.cuh"
#include "../../include/mtvaf_b200.h"

namespace mtvaf {

constexpr int LN_MAX_VEC = 4;           // 4 vectors x 8 elements x 32 lanes = H <= 1024
constexpr int LN_WARPS = 4;

template <typename T>
__device__ __forceinline__ void ln_load_row(const T* row, int H, int lane, float (&x)[LN_MAX_VEC][8]) {
#pragma unroll
  for (int v = 0; v < LN_MAX_VEC; ++v) {
    const int c = (v * 32 + lane) * 8;
    if (c < H) Vec8<T>::load(row + c, x[v]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[v][j] = 0.f;
    }
  }
}

__device__ __forceinline__ void ln_stats(const float (&x)[LN_MAX_VEC][8], int H, int lane, float eps, float& mean,
                                         float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < LN_MAX_VEC; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[v][j];
  mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int v = 0; v < LN_MAX_VEC; ++v) {
    const int c = (v * 32 + lane) * 8;
    if (c < H) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = x[v][j] - mean; q += d * d; }
    }
  }
  rstd = rsqrtf(warp_sum(q) / (float)H + eps);
}

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const T* __restrict__ z, T* __restrict__ y, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, int rows, int H, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  float x[LN_MAX_VEC][8];
  ln_load_row<T>(z + (long long)row * H, H, lane, x);
  float mean, rstd;
  ln_stats(x, H, lane, eps, mean, rstd);
#pragma unroll
  for (int v = 0; v < LN_MAX_VEC; ++v) {
    const int c = (v * 32 + lane) * 8;
    if (c < H) {
      float g[8], b[8], o[8];
      Vec8<float>::load(gamma + c, g);
      Vec8<float>::load(beta + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (x[v][j] - mean) * rstd * g[j] + b[j];
      Vec8<T>::store(y + (long long)row * H + c, o);
    }
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  xhat = (z - mean) * rstd
// d_gamma += sum_rows dy * xhat, d_beta += sum_rows dy (per-warp register partials -> smem -> atomics).
// Fused tail of the backward of `LN(dropout(dense) + residual)` (modeling_roberta.py:296-298,379-381):
//   dd = dropout_mask o dz / (1-p)  (the gradient entering the dense GEMMs; written when p > 0),
//   d_bias += sum_rows dd           (bias gradient of that dense layer),
// so the separate dropout and column-sum passes over [T,H] disappear.
// NV = ceil(H / 256) vectors of 8 elements per lane; the next row's loads are issued before the current
// row's reductions (raw 16-byte vectors held in registers) to keep enough bytes in flight per SM.
// (A cp.async ring in shared memory with three rows in flight per warp was measured SLOWER, 51.5 vs 45.1 us: the
// kernel is issue-bound -- 886 warp instructions per row, 56 % issue utilisation -- not latency-bound.)
// one row of the backward: x / g hold z and dy on entry
template <typename T, int NV>
__device__ __forceinline__ void ln_bwd_row(float (&x)[NV][8], float (&g)[NV][8], float mean, float rstd, int row, int H,
                                           int lane, const float* __restrict__ gamma, T* __restrict__ dz,
                                           T* __restrict__ dd, bool want_bias, uint32_t drop_thr, float drop_scale,
                                           unsigned long long seed, float (&ag)[NV][8], float (&ab)[NV][8],
                                           float (&ac)[NV][8]) {
  using V = Vec8<T>;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * 32 + lane) * 8;
    if (c < H) {
      float gm[8];
      Vec8<float>::load(gamma + c, gm);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[v][j] - mean) * rstd;
        const float d = g[v][j];
        ag[v][j] += d * xh;
        ab[v][j] += d;
        const float gg = d * gm[j];
        x[v][j] = xh;
        g[v][j] = gg;
        s1 += gg;
        s2 += gg * xh;
      }
    }
  }
  s1 = warp_sum(s1) / (float)H;
  s2 = warp_sum(s2) / (float)H;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = (v * 32 + lane) * 8;
    if (c < H) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[v][j] - s1 - x[v][j] * s2);
      V::store(dz + (long long)row * H + c, o);
      if (drop_thr) {
        const uint32_t keep = dropout_keep8(seed, (unsigned long long)row * H + c, drop_thr);   // H % 8 == 0
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = ((keep >> j) & 1u) ? o[j] * drop_scale : 0.f;
        V::store(dd + (long long)row * H + c, o);
      }
      if (want_bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) ac[v][j] += o[j];
      }
    }
  }
}

// reduce the per-warp parameter-gradient partials across the block, then one atomic per column
template <int NV>
__device__ __forceinline__ void ln_bwd_flush(float* red, int H, const float (&ag)[NV][8], const float (&ab)[NV][8],
                                             const float (&ac)[NV][8], float* d_gamma, float* d_beta, float* d_bias) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int pass = 0; pass < (d_bias ? 3 : 2); ++pass) {
    float* dst = pass == 0 ? d_gamma : pass == 1 ? d_beta : d_bias;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp * 256 + lane * 8 + j] = pass == 0 ? ag[v][j] : pass == 1 ? ab[v][j] : ac[v][j];
      __syncthreads();
      for (int i = threadIdx.x; i < 256; i += LN_WARPS * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) s += red[w * 256 + i];
        const int c = v * 256 + i;
        if (c < H) atomicAdd(dst + c, s);
      }
    }
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, (NV <= 3 && sizeof(T) == 2) ? 3 : 2)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ z, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, int rows, int H,
                     T* __restrict__ dz, T* __restrict__ dd, float* __restrict__ d_gamma, float* __restrict__ d_beta,
                     float* __restrict__ d_bias, uint32_t drop_thr, float drop_scale, unsigned long long seed_in,
                     const unsigned long long* __restrict__ step) {
  __shared__ float red[LN_WARPS * 256];
  using V = Vec8<T>;
  const unsigned long long seed = drop_thr ? step_seed(seed_in, step) : seed_in;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ag[NV][8], ab[NV][8], ac[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[v][j] = 0.f; ab[v][j] = 0.f; ac[v][j] = 0.f; }
  const int stride = gridDim.x * LN_WARPS;
  int row = blockIdx.x * LN_WARPS + warp;
  typename V::Raw rz[NV], rg[NV];
  float mean_n = 0.f, rstd_n = 0.f;              // the row statistics travel with the row (they were 24 % of all
  if (row < rows) {                              // stall samples when loaded at first use, profiles/r1_ncu_hot_v7.md)
    mean_n = __ldg(mean_in + row);
    rstd_n = __ldg(rstd_in + row);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 8;
      if (c < H) { rz[v] = V::load_raw(z + (long long)row * H + c); rg[v] = V::load_raw(dy + (long long)row * H + c); }
    }
  }
  for (; row < rows; row += stride) {
    float x[NV][8], g[NV][8];
#pragma unroll
    for (int v = 0; v < NV; ++v) { V::cvt(rz[v], x[v]); V::cvt(rg[v], g[v]); }
    const float mean = mean_n, rstd = rstd_n;
    const int nrow = row + stride;
    if (nrow < rows) {
      mean_n = __ldg(mean_in + nrow);
      rstd_n = __ldg(rstd_in + nrow);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = (v * 32 + lane) * 8;
        if (c < H) { rz[v] = V::load_raw(z + (long long)nrow * H + c); rg[v] = V::load_raw(dy + (long long)nrow * H + c); }
      }
    }
    ln_bwd_row<T, NV>(x, g, mean, rstd, row, H, lane, gamma, dz, dd, d_bias != nullptr, drop_thr, drop_scale, seed, ag, ab, ac);
  }
  ln_bwd_flush<NV>(red, H, ag, ab, ac, d_gamma, d_beta, d_bias);
}


// ------------------------------------------------------------------ position ids (bit-exact int64)
// roberta: mask = ids != pad; pos = cumsum(mask) * mask + pad   (modeling_roberta.py:1717-1719)
__global__ void position_ids_kernel(const int64_t* __restrict__ ids, int64_t* __restrict__ pos, int B, int L,
                                    int kind, int pad) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  int running = 0;
  for (int base = 0; base < L; base += 32) {
    const int i = base + lane;
    if (kind == 1) {
      if (i < L) pos[(long long)b * L + i] = i;
      continue;
    }
    const int m = (i < L) ? (ids[(long long)b * L + i] != pad) : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    const int incl = __popc(bal & (0xffffffffu >> (31 - lane)));
    if (i < L) pos[(long long)b * L + i] = (int64_t)((running + incl) * m) + pad;
    running += __popc(bal);
  }
}

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32)
embed_ln_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tts, const int64_t* __restrict__ pos,
                    const float* __restrict__ word, const float* __restrict__ pemb, const float* __restrict__ temb,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int rows, int H,
                    T* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                    uint32_t drop_thr, float drop_scale, unsigned long long seed_in,
                    const unsigned long long* __restrict__ step) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const unsigned long long seed = drop_thr ? step_seed(seed_in, step) : seed_in;
  const float* w = word + ids[row] * (long long)H;
  const float* p = pemb + pos[row] * (long long)H;
  const float* t = temb + tts[row] * (long long)H;
  float x[LN_MAX_VEC][8];
#pragma unroll
  for (int v = 0; v < LN_MAX_VEC; ++v) {
    const int c = (v * 32 + lane) * 8;
    if (c < H) {
      float a[8], b[8], d[8];
      Vec8<float>::load(w + c, a);
      Vec8<float>::load(t + c, b);
      Vec8<float>::load(p + c, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[v][j] = (a[j] + b[j]) + d[j];    // same association as :134-137
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[v][j] = 0.f;
    }
  }
  float mean, rstd;
  ln_stats(x, H, lane, eps, mean, rstd);
#pragma unroll
  for (int v = 0; v < LN_MAX_VEC; ++v) {
    const int c = (v * 32 + lane) * 8;
    if (c < H) {
      float g[8], b[8], o[8];
      Vec8<float>::load(gamma + c, g);
      Vec8<float>::load(beta + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = (x[v][j] - mean) * rstd * g[j] + b[j];
        if (drop_thr) o[j] = dropout_keep(seed, (unsigned long long)row * H + c + j, drop_thr) ? o[j] * drop_scale : 0.f;
      }
      Vec8<T>::store(out + (long long)row * H + c, o);
    }
  }
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// Backward of the embedding sum + LayerNorm.  The three scatter-adds are the hard part: every token adds a
// full row to word[id], pos[pid] and type[tt], and on real batches most tokens collide (all token types are 0,
// each position id occurs once per sequence, ~3/4 of the ids are the padding token), so per-element atomics
// serialise in L2.  Each warp therefore walks the tokens in POSITION-major order (l outer, b inner: the
// position row and the type row stay the same for a whole run) and keeps register accumulators
//   ade : sum of dE over the current run of equal (token type, position id)   -> flushed to both tables
//   aw  : sum of dE over the rows whose id equals the batch's hot id (the padding token) -> flushed once
// next to the d_gamma / d_beta partials; only rows with other ids scatter directly (16-byte vector atomics).
__device__ __forceinline__ void red_add_v8(float* p, const float (&v)[8]) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  atomicAdd(reinterpret_cast<float4*>(p + 4), make_float4(v[4], v[5], v[6], v[7]));
}

template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
embed_ln_bwd_kernel(const T* __restrict__ dout, const int64_t* __restrict__ ids, const int64_t* __restrict__ tts,
                    const int64_t* __restrict__ pos, const float* __restrict__ word, const float* __restrict__ pemb,
                    const float* __restrict__ temb, const float* __restrict__ gamma, const float* __restrict__ mean_in,
                    const float* __restrict__ rstd_in, int B, int L, int H, int word_pad, int pos_pad,
                    float* __restrict__ d_word, float* __restrict__ d_pos, float* __restrict__ d_type,
                    float* __restrict__ d_gamma, float* __restrict__ d_beta, uint32_t drop_thr, float drop_scale,
                    unsigned long long seed_in, const unsigned long long* __restrict__ step) {
  __shared__ float red[LN_WARPS * 256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = B * L;
  const unsigned long long seed = drop_thr ? step_seed(seed_in, step) : seed_in;
  float ag[NV][8], ab[NV][8], ade[NV][8], aw[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[v][j] = 0.f; ab[v][j] = 0.f; ade[v][j] = 0.f; aw[v][j] = 0.f; }
  const long long hot = ids[rows - 1];            // the last token of the batch: the padding id if there is any padding
  long long run_tt = -1, run_ps = -1;
  auto flush_run = [&]() {
    if (run_tt < 0) return;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 8;
      if (c < H) {
        red_add_v8(d_type + run_tt * (long long)H + c, ade[v]);
        if (run_ps != pos_pad) red_add_v8(d_pos + run_ps * (long long)H + c, ade[v]);
#pragma unroll
        for (int j = 0; j < 8; ++j) ade[v][j] = 0.f;
      }
    }
  };
  // contiguous chunk of the position-major token order per warp
  const int n_warps = gridDim.x * LN_WARPS;
  const int chunk = (rows + n_warps - 1) / n_warps;
  const int j0 = (blockIdx.x * LN_WARPS + warp) * chunk;
  const int j1 = min(rows, j0 + chunk);
  for (int jj = j0; jj < j1; ++jj) {
    const int l = jj / B, bb = jj - l * B;
    const int row = bb * L + l;
    const long long id = ids[row], ps = pos[row], tt = tts[row];
    if (tt != run_tt || ps != run_ps) { flush_run(); run_tt = tt; run_ps = ps; }
    const float* w = word + id * (long long)H;
    const float* p = pemb + ps * (long long)H;
    const float* t = temb + tt * (long long)H;
    float x[NV][8], g[NV][8];
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 8;
      if (c < H) {
        float a[8], b[8], d[8], gm[8];
        Vec8<T>::load(dout + (long long)row * H + c, g[v]);
        Vec8<float>::load(w + c, a);
        Vec8<float>::load(t + c, b);
        Vec8<float>::load(p + c, d);
        Vec8<float>::load(gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float dd = g[v][j];
          if (drop_thr) dd = dropout_keep(seed, (unsigned long long)row * H + c + j, drop_thr) ? dd * drop_scale : 0.f;
          const float xh = (((a[j] + b[j]) + d[j]) - mean) * rstd;
          ag[v][j] += dd * xh;
          ab[v][j] += dd;
          const float gg = dd * gm[j];
          x[v][j] = xh;
          g[v][j] = gg;
          s1 += gg;
          s2 += gg * xh;
        }
      }
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
    const bool is_hot = id == hot;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 8;
      if (c < H) {
        float de[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          de[j] = rstd * (g[v][j] - s1 - x[v][j] * s2);
          ade[v][j] += de[j];
          if (is_hot) aw[v][j] += de[j];
        }
        if (!is_hot && id != word_pad) red_add_v8(d_word + id * (long long)H + c, de);
      }
    }
  }
  flush_run();
  if (hot != word_pad && j1 > j0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * 32 + lane) * 8;
      if (c < H) red_add_v8(d_word + hot * (long long)H + c, aw[v]);
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp * 256 + lane * 8 + j] = pass == 0 ? ag[v][j] : ab[v][j];
      __syncthreads();
      for (int i = threadIdx.x; i < 256; i += LN_WARPS * 32) {
        float s = 0.f;
#pragma unroll
        for (int ww = 0; ww < LN_WARPS; ++ww) s += red[ww * 256 + i];
        const int c = v * 256 + i;
        if (c < H) atomicAdd((pass == 0 ? d_gamma : d_beta) + c, s);
      }
    }
  }
}

static int drop_params(float p, uint32_t* thr, float* scale) {
  *thr = 0; *scale = 1.f;
  if (p > 0.f) {
    MTVAF_REQUIRE(p < 1.f, "dropout p must be < 1");
    double t = (double)p * 4294967296.0;
    *thr = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    *scale = 1.f / (1.f - p);
  }
  return 0;
}

}  // namespace mtvaf

using namespace mtvaf;

extern "C" int mtvaf_layernorm_fwd(const void* z, void* y, const float* gamma, const float* beta, float eps,
                                   int rows, int H, int dtype, float* mean, float* rstd, void* stream) {
  MTVAF_REQUIRE(z && y && gamma && beta && rows > 0, "layernorm_fwd: bad argument");
  MTVAF_REQUIRE(H % 8 == 0 && H <= LN_MAX_VEC * 256, "layernorm: H=%d must be a multiple of 8 and <= %d", H,
                LN_MAX_VEC * 256);
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  if (dtype == MTVAF_BF16)
    layernorm_fwd_kernel<__nv_bfloat16><<<grid, LN_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)z, (__nv_bfloat16*)y, gamma, beta, eps, rows, H, mean, rstd);
  else
    layernorm_fwd_kernel<float><<<grid, LN_WARPS * 32, 0, (cudaStream_t)stream>>>((const float*)z, (float*)y, gamma,
                                                                                 beta, eps, rows, H, mean, rstd);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_ln_bwd(const void* dy, const void* z, const float* gamma, const float* mean, const float* rstd,
                         int rows, int H, void* dz, void* dd, float* d_gamma, float* d_beta, float* d_bias,
                         uint32_t thr, float scale, uint64_t seed, cudaStream_t st) {
  int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  const int nv = (H + 255) / 256;
  const int cap = sm_count() * ((nv <= 3 && sizeof(T) == 2) ? 3 : 2);   // resident blocks per SM (register-limited)
  if (grid > cap) grid = cap;
#define MTVAF_LN_BWD(NV_)                                                                                          \
  layernorm_bwd_kernel<T, NV_><<<grid, LN_WARPS * 32, 0, st>>>((const T*)dy, (const T*)z, gamma, mean, rstd, rows, H, \
                                                              (T*)dz, (T*)dd, d_gamma, d_beta, d_bias, thr, scale, seed, \
                                                              step_source())
  if (nv == 1) MTVAF_LN_BWD(1);
  else if (nv == 2) MTVAF_LN_BWD(2);
  else if (nv == 3) MTVAF_LN_BWD(3);
  else MTVAF_LN_BWD(4);
#undef MTVAF_LN_BWD
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_layernorm_bwd(const void* dy, const void* z, const float* gamma, const float* mean,
                                   const float* rstd, int rows, int H, int dtype, void* dz, float* d_gamma,
                                   float* d_beta, void* dd, float* d_bias, float p_drop, uint64_t seed,
                                   void* stream) {
  MTVAF_REQUIRE(dy && z && gamma && mean && rstd && dz && d_gamma && d_beta && rows > 0, "layernorm_bwd: bad argument");
  MTVAF_REQUIRE(H % 8 == 0 && H <= LN_MAX_VEC * 256, "layernorm: H=%d unsupported", H);
  uint32_t thr; float scale;
  if (int rc = drop_params(p_drop, &thr, &scale)) return rc;
  MTVAF_REQUIRE(thr == 0 || dd != nullptr, "layernorm_bwd: p_drop > 0 needs the `dd` output");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTVAF_BF16)
    return launch_ln_bwd<__nv_bfloat16>(dy, z, gamma, mean, rstd, rows, H, dz, dd, d_gamma, d_beta, d_bias, thr,
                                        scale, seed, st);
  return launch_ln_bwd<float>(dy, z, gamma, mean, rstd, rows, H, dz, dd, d_gamma, d_beta, d_bias, thr, scale, seed, st);
}

extern "C" int mtvaf_embed_ln_fwd(const int64_t* input_ids, const int64_t* token_type_ids, const float* word_emb,
                                  const float* pos_emb, const float* type_emb, const float* gamma, const float* beta,
                                  float eps, int kind, int pad_idx, int B, int L, int H, int vocab, int max_pos,
                                  int n_types, void* out, int out_dtype, int64_t* position_ids, float* mean,
                                  float* rstd, float p_drop, uint64_t seed, void* stream) {
  MTVAF_REQUIRE(input_ids && token_type_ids && word_emb && pos_emb && type_emb && gamma && beta && out &&
                    position_ids && mean && rstd, "embed_ln_fwd: null argument");
  MTVAF_REQUIRE(B > 0 && L > 0 && H % 8 == 0 && H <= LN_MAX_VEC * 256, "embed_ln_fwd: bad shape");
  // position ids can reach L + pad_idx (roberta): the table must hold them (514 for L <= 512)
  MTVAF_REQUIRE((kind == 0 ? L + pad_idx : L - 1) < max_pos, "embed_ln_fwd: sequence length %d exceeds the %d learned positions", L, max_pos);
  (void)vocab; (void)n_types;
  uint32_t thr; float scale;
  if (int rc = drop_params(p_drop, &thr, &scale)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  position_ids_kernel<<<(B + 3) / 4, 128, 0, st>>>(input_ids, position_ids, B, L, kind, pad_idx);
  MTVAF_LAUNCH_CHECK();
  const int rows = B * L;
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  if (out_dtype == MTVAF_BF16)
    embed_ln_fwd_kernel<__nv_bfloat16><<<grid, LN_WARPS * 32, 0, st>>>(input_ids, token_type_ids, position_ids,
                                                                      word_emb, pos_emb, type_emb, gamma, beta, eps,
                                                                      rows, H, (__nv_bfloat16*)out, mean, rstd, thr,
                                                                      scale, seed, step_source());
  else
    embed_ln_fwd_kernel<float><<<grid, LN_WARPS * 32, 0, st>>>(input_ids, token_type_ids, position_ids, word_emb,
                                                              pos_emb, type_emb, gamma, beta, eps, rows, H,
                                                              (float*)out, mean, rstd, thr, scale, seed, step_source());
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_embed_ln_bwd(const void* dout, int dtype, const int64_t* input_ids,
                                  const int64_t* token_type_ids, const int64_t* position_ids, const float* word_emb,
                                  const float* pos_emb, const float* type_emb, const float* gamma, const float* mean,
                                  const float* rstd, int kind, int pad_idx, int B, int L, int H, float* d_word,
                                  float* d_pos, float* d_type, float* d_gamma, float* d_beta, float p_drop,
                                  uint64_t seed, void* stream) {
  MTVAF_REQUIRE(dout && input_ids && token_type_ids && position_ids && word_emb && pos_emb && type_emb && gamma &&
                    mean && rstd && d_word && d_pos && d_type && d_gamma && d_beta, "embed_ln_bwd: null argument");
  MTVAF_REQUIRE(H % 8 == 0 && H <= LN_MAX_VEC * 256, "embed_ln_bwd: bad H");
  uint32_t thr; float scale;
  if (int rc = drop_params(p_drop, &thr, &scale)) return rc;
  const int rows = B * L;
  int grid = (rows + LN_WARPS * 8 - 1) / (LN_WARPS * 8);      // >= 8 tokens per warp so runs can aggregate
  const int cap = sm_count() * 2;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  // nn.Embedding(padding_idx): word table always; position table only for roberta (modeling_roberta.py:98-100)
  const int word_pad = pad_idx, pos_pad = (kind == 0) ? pad_idx : -1;
  const int nv = (H + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
#define MTVAF_EMB_BWD(T_, NV_)                                                                                      \
  embed_ln_bwd_kernel<T_, NV_><<<grid, LN_WARPS * 32, 0, st>>>((const T_*)dout, input_ids, token_type_ids,           \
      position_ids, word_emb, pos_emb, type_emb, gamma, mean, rstd, B, L, H, word_pad, pos_pad, d_word, d_pos,      \
      d_type, d_gamma, d_beta, thr, scale, seed, step_source())
  if (dtype == MTVAF_BF16) {
    if (nv == 1) MTVAF_EMB_BWD(__nv_bfloat16, 1); else if (nv == 2) MTVAF_EMB_BWD(__nv_bfloat16, 2);
    else if (nv == 3) MTVAF_EMB_BWD(__nv_bfloat16, 3); else MTVAF_EMB_BWD(__nv_bfloat16, 4);
  } else {
    if (nv == 1) MTVAF_EMB_BWD(float, 1); else if (nv == 2) MTVAF_EMB_BWD(float, 2);
    else if (nv == 3) MTVAF_EMB_BWD(float, 3); else MTVAF_EMB_BWD(float, 4);
  }
#undef MTVAF_EMB_BWD
  MTVAF_LAUNCH_CHECK();
  return 0;
}
