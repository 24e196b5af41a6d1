// Shared device helpers: error handling, bf16 packing, warp/block reductions, counter-based dropout RNG.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mtvaf {

// ---- error reporting across the C ABI (never throw) -------------------------------------------
void set_last_error(const char* fmt, ...);
#define MTVAF_CHECK_CUDA(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::mtvaf::set_last_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,              \
                              cudaGetErrorName(_e), cudaGetErrorString(_e));              \
      return -2;                                                                          \
    }                                                                                     \
  } while (0)
#define MTVAF_REQUIRE(cond, ...)                                                          \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::mtvaf::set_last_error(__VA_ARGS__);                                               \
      return -1;                                                                          \
    }                                                                                     \
  } while (0)
// every kernel launch of the library goes through this check; it also counts launches (mtvaf_launch_count)
void note_launch();
#define MTVAF_LAUNCH_CHECK()              \
  do {                                    \
    ::mtvaf::note_launch();               \
    MTVAF_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

int sm_count();
// device pointer registered with mtvaf_set_step_source (or NULL): a step counter in HBM that every dropout site
// mixes into its seed, so a CUDA graph captured once draws fresh masks on every replay
const unsigned long long* step_source();

// ---- element type helpers -----------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// 8-element vector load/store (16 B for bf16, 2x16 B for fp32); pointers must be aligned accordingly
template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw load_raw(const float* p) {
    Raw r;
    r.a = *reinterpret_cast<const float4*>(p);
    r.b = *reinterpret_cast<const float4*>(p + 4);
    return r;
  }
  static __device__ __forceinline__ void cvt(const Raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct Vec8<__nv_bfloat16> {
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void cvt(const Raw& u, float (&v)[8]) {
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

// ---- reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// ---- math ---------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7): one MUFU.RCP + one MUFU.EX2 + ~12 FMA-class ops,
// used by the bf16 tensor-core epilogues (results are rounded to bf16, eps 3.9e-3); the fp32 parity path
// keeps erff.  `e_out` = exp(-x^2/2), shared with the Gaussian pdf of GELU'.
__device__ __forceinline__ float normal_cdf_fast(float x, float& e_out) {
  // t = 1 / (1 + p |x|/sqrt2);  e = exp(-x^2/2);  0.5 erfc(|x|/sqrt2) = t (c1 + t (c2 + ...)) e  (c_i = a_i / 2)
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752f, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * -0.72134752044448170f) * x));
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(t, p, 0.5f * 1.421413741f);
  p = fmaf(t, p, 0.5f * -0.284496736f);
  p = fmaf(t, p, 0.5f * 0.254829592f);
  e_out = e;
  const float h = (p * t) * e;
  return x >= 0.f ? 1.f - h : h;
}
// GELU for the bf16 tensor-core epilogues, which are issue-bound (20 instructions per element do not hide behind
// a K = 768 mainloop): the tanh form 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) on MUFU.TANH, 6 instructions.
// It deviates from the erf form of the reference (ACT2FN["gelu"], modeling_roberta.py:366) by at most 4.7e-4
// absolute (8.7e-4 for the derivative) -- below bf16 resolution of O(1) activations and 40x inside the 2e-2
// bf16 tolerance; forward and backward use the same function and its exact derivative.  The fp32 parity path
// keeps erff (gelu_erf / dgelu_erf above).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_fast(float x) {
#ifdef MTVAF_GELU_EXACT      // experiment builds (tools/bf16_error_budget.py): how much of the bf16 error is the tanh form?
  return gelu_erf(x);
#endif
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(x2, 0.0356774081f, 0.7978845608f));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ float dgelu_fast(float x) {
#ifdef MTVAF_GELU_EXACT
  return dgelu_erf(x);
#endif
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(x2, 0.0356774081f, 0.7978845608f));
  const float du = fmaf(x2, 0.1070322243f, 0.7978845608f);      // d/dx of the tanh argument
  const float hs = (0.5f * x) * fmaf(-t, t, 1.f);                // 0.5 x sech^2
  return fmaf(hs, du, fmaf(0.5f, t, 0.5f));
}

// gelu and its derivative from ONE tanh (MTVAF_EPI_GELU_GRAD: the derivative is saved by the forward epilogue)
__device__ __forceinline__ void gelu_and_grad_fast(float x, float& g, float& d) {
#ifdef MTVAF_GELU_EXACT
  g = gelu_erf(x);
  d = dgelu_erf(x);
  return;
#endif
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(x2, 0.0356774081f, 0.7978845608f));
  const float hx = 0.5f * x;
  g = fmaf(hx, t, hx);
  const float du = fmaf(x2, 0.1070322243f, 0.7978845608f);
  d = fmaf(hx * fmaf(-t, t, 1.f), du, fmaf(0.5f, t, 0.5f));
}

// ---- dropout: counter-based hash RNG, recomputed (never stored) in backward ----------------------
// keep(seed, stream, idx) is a pure function; `stream` separates the dropout sites of one step.
__device__ __forceinline__ uint32_t dropout_bits(uint64_t seed, uint64_t idx) {
  // two multiply rounds over (idx, seed) with a xor-shift in between: 8 integer ops per element (the GEMM
  // epilogues and the attention softmax are issue-bound, so the mask must be cheap).  Only the HIGH bits
  // matter (the caller compares against a threshold), and those are fully mixed by the second multiply.
  const uint32_t lo = static_cast<uint32_t>(idx), hi = static_cast<uint32_t>(idx >> 32);
  const uint32_t s0 = static_cast<uint32_t>(seed), s1 = static_cast<uint32_t>(seed >> 32);
  uint32_t h = (lo ^ s0) * 0x9E3779B1u;
  h ^= h >> 15;
  h = (h ^ (hi * 0x85EBCA77u + s1)) * 0xC2B2AE3Du;
  return h;
}
// effective seed of a dropout site: the by-value seed plus the device-resident step counter (if registered)
__device__ __forceinline__ unsigned long long step_seed(unsigned long long seed, const unsigned long long* step) {
  return step ? seed + (*step) * 0x9E3779B97F4A7C15ull : seed;
}
// Hidden-state dropout mask: ONE hash per aligned QUAD of elements (idx >> 2), whose 32 x 64 -> 64-bit product is
// consumed as four fields (lo, lo << 16, hi, hi << 16), each compared with the full 32-bit threshold
// threshold = round(p * 2^32): fields 0 / 2 are exact to 2^-32, fields 1 / 3 have 16-bit resolution (|dp| <= 1.6e-5).
// Element idx is KEPT iff field (idx & 3) of quad (idx >> 2) >= threshold.  ~2.5 integer ops per element instead of
// 9: the RESID GEMM epilogue and the LayerNorm backward are issue-bound (profiles/r1_ncu_hot_v7.md).
__device__ __forceinline__ void dropout_quad(uint64_t seed, uint64_t quad, uint32_t (&f)[4]) {
  const uint32_t lo = static_cast<uint32_t>(quad), hi = static_cast<uint32_t>(quad >> 32);
  const uint32_t s0 = static_cast<uint32_t>(seed), s1 = static_cast<uint32_t>(seed >> 32);
  uint32_t h = (lo ^ s0) * 0x9E3779B1u;
  h ^= h >> 15;
  h ^= hi * 0x85EBCA77u + s1;
  const unsigned long long w = static_cast<unsigned long long>(h) * 0x85EBCA77C2B2AE3Dull;   // IMAD.WIDE + IMAD
  const uint32_t wl = static_cast<uint32_t>(w), wh = static_cast<uint32_t>(w >> 32);
  f[0] = wl;
  f[1] = wl << 16;
  f[2] = wh;
  f[3] = wh << 16;
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, uint32_t threshold) {
  uint32_t f[4];
  dropout_quad(seed, idx >> 2, f);
  const uint32_t i = static_cast<uint32_t>(idx) & 3u;
  return (i == 0 ? f[0] : i == 1 ? f[1] : i == 2 ? f[2] : f[3]) >= threshold;
}
// keep bits of the 8 consecutive elements idx0 .. idx0 + 7 (bit j = element idx0 + j)
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t idx0, uint32_t threshold) {
  uint32_t keep = 0;
  if ((static_cast<uint32_t>(idx0) & 3u) == 0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t f[4];
      dropout_quad(seed, (idx0 >> 2) + q, f);
#pragma unroll
      for (int j = 0; j < 4; ++j) keep |= (f[j] >= threshold ? 1u : 0u) << (4 * q + j);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) keep |= (dropout_keep(seed, idx0 + j, threshold) ? 1u : 0u) << j;
  }
  return keep;
}

// ---- attention-probability dropout (modeling_roberta.py:268) -----------------------------------
// The softmax kernels are issue-bound (profiles/r1_ncu_hot_v7.md: 21 instructions per element in the backward
// loop, 9 of them mask arithmetic), so the mask costs < 3 integer ops per element: one 32-bit key per
// (batch, head, query) row (attn_drop_rowkey, hashed once per row) and ONE hash per QUAD of keys whose
// 32x64 -> 64-bit product is consumed as four fields: lo, lo << 16, hi, hi << 16, each compared with the full
// 32-bit threshold (fields 0 / 2 are exact to 2^-32, fields 1 / 3 have 16-bit resolution: |dp| <= 1.6e-5).
// Reference key numbering: prefix rows 0..P-1, then text rows.  Forward, backward and the SIMT kernels all
// derive the mask from these functions, so backward regenerates forward's mask.
__device__ __forceinline__ uint32_t attn_drop_rowkey(uint64_t seed, uint64_t row) { return dropout_bits(seed, row); }
__device__ __forceinline__ void attn_drop_quad(uint32_t rowkey, uint32_t quad, uint32_t (&f)[4]) {
  uint32_t h = (quad ^ rowkey) * 0x9E3779B1u;
  h ^= h >> 15;
  // low 64 bits of h x (64-bit odd constant): IMAD.WIDE.U32 + IMAD.  A 32-bit multiplier would leave the high word
  // in [0, C) instead of [0, 2^32) (measured: keep rate 0.868 instead of 0.9 for that field)
  const unsigned long long w = static_cast<unsigned long long>(h) * 0x85EBCA77C2B2AE3Dull;
  const uint32_t lo = static_cast<uint32_t>(w), hi = static_cast<uint32_t>(w >> 32);
  f[0] = lo;
  f[1] = lo << 16;
  f[2] = hi;
  f[3] = hi << 16;
}
__device__ __forceinline__ bool attn_drop_keep(uint32_t rowkey, int kk, uint32_t thr) {
  uint32_t f[4];
  attn_drop_quad(rowkey, static_cast<uint32_t>(kk) >> 2, f);
  const int i = kk & 3;
  return (i == 0 ? f[0] : i == 1 ? f[1] : i == 2 ? f[2] : f[3]) >= thr;
}
// zero the dropped entries among the 8 consecutive keys kk0 .. kk0+7 (kk0 >= 0) of v (and w): the fields of a quad
// are consumed as predicates right away (no flag array in registers); the 1/(1-p) scale is the caller's business
__device__ __forceinline__ void attn_drop_apply8(uint32_t rowkey, int kk0, uint32_t thr, float (&v)[8]) {
  if ((kk0 & 3) == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint32_t f[4];
      attn_drop_quad(rowkey, static_cast<uint32_t>(kk0 >> 2) + i, f);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (f[j] < thr) v[4 * i + j] = 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!attn_drop_keep(rowkey, kk0 + j, thr)) v[j] = 0.f;
  }
}
__device__ __forceinline__ void attn_drop_apply8(uint32_t rowkey, int kk0, uint32_t thr, float (&v)[8], float (&w)[8]) {
  if ((kk0 & 3) == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint32_t f[4];
      attn_drop_quad(rowkey, static_cast<uint32_t>(kk0 >> 2) + i, f);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (f[j] < thr) { v[4 * i + j] = 0.f; w[4 * i + j] = 0.f; }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!attn_drop_keep(rowkey, kk0 + j, thr)) { v[j] = 0.f; w[j] = 0.f; }
  }
}

}  // namespace mtvaf
