// Coalesced GEMM epilogue for the CTA-pair tcgen05 kernel: every epilogue warp owns a [32 rows x 64 cols]
// bf16 box of the output tile at a time.  The residual / pre-activation operand of the fused epilogues
// (models/modeling_roberta.py:297-298,366 and their backward) arrives by TMA into a per-warp
// SWIZZLE_128B box (prefetched one box ahead; shared memory goes to the operand ring first), the results are written to a per-warp
// staging box in shared memory (conflict-free 16-byte pieces) and leave with ONE TMA store per box, so
// global traffic is full 128-byte lines instead of 32 scattered 16-byte pieces per warp instruction.
#pragma once
#include "epilogue.cuh"
#include "ptx.cuh"

namespace mtvaf {

constexpr int kEpiBoxBytes = 32 * 128;     // 32 rows x 64 bf16

template <int MODE>
struct StagedEpi {
  static constexpr bool kStaged = (MODE == MTVAF_EPI_STORE || MODE == MTVAF_EPI_GELU || MODE == MTVAF_EPI_TANH ||
                                   MODE == MTVAF_EPI_RESID || MODE == MTVAF_EPI_MUL_DGELU ||
                                   MODE == MTVAF_EPI_MUL_DTANH || MODE == MTVAF_EPI_GELU_GRAD ||
                                   MODE == MTVAF_EPI_MUL_AUX);
  static constexpr bool kAux = (MODE == MTVAF_EPI_RESID || MODE == MTVAF_EPI_MUL_DGELU || MODE == MTVAF_EPI_MUL_DTANH ||
                                MODE == MTVAF_EPI_MUL_AUX);
  static constexpr bool kTwoOut = (MODE == MTVAF_EPI_GELU || MODE == MTVAF_EPI_GELU_GRAD);   // out + out2 boxes
#ifndef MTVAF_EPI_INPLACE
#define MTVAF_EPI_INPLACE 0                // 1: the aux variants write their results INTO the aux box and store from it
#endif                                     //    (no separate staging box: one more operand-ring stage; the next aux load waits for the store)
  static constexpr bool kInplace = kAux && (MTVAF_EPI_INPLACE != 0);
  static constexpr int kOutBufs = kStaged ? (kInplace ? 0 : (kTwoOut ? 2 : 1)) : 0;
#ifndef MTVAF_EPI_AUX_BUFS
// aux boxes per warp pair: 1 = the next box's operand is requested as soon as both warps have read the current one
// (the operand ring gets the 32 KB back: 5 stages instead of 4 for RESID / xGELU' / xtanh'), 2 = prefetched two boxes
// ahead.  Measured in isolation at M = 65536 (tools/bench_gemm.py, round 2): 1 is 5-14 % faster on every fused-aux variant.
#define MTVAF_EPI_AUX_BUFS 1
#endif
  static constexpr int kAuxBufs = kAux ? MTVAF_EPI_AUX_BUFS : 0;
  static constexpr int kBytesPerWarp = (kOutBufs + kAuxBufs) * kEpiBoxBytes;
};

// dropout keep-mask bits for 32 consecutive elements starting at 64-bit index idx0 (same function as
// dropout_keep(seed, idx0 + j, thr) of common.cuh): eight quads, seed / high-word terms hoisted out of the loop
__device__ __forceinline__ uint32_t dropout_mask32(unsigned long long seed, unsigned long long idx0, uint32_t thr) {
  uint32_t keep = 0;
  const unsigned long long q0 = idx0 >> 2;
  if ((static_cast<uint32_t>(idx0) & 3u) == 0 && static_cast<uint32_t>(q0) <= 0xFFFFFFFFu - 7u) {
    const uint32_t s0 = static_cast<uint32_t>(seed), s1 = static_cast<uint32_t>(seed >> 32);
    const uint32_t lo0 = static_cast<uint32_t>(q0);
    const uint32_t hterm = static_cast<uint32_t>(q0 >> 32) * 0x85EBCA77u + s1;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t h = ((lo0 + q) ^ s0) * 0x9E3779B1u;
      h ^= h >> 15;
      h ^= hterm;
      const unsigned long long w = static_cast<unsigned long long>(h) * 0x85EBCA77C2B2AE3Dull;
      const uint32_t wl = static_cast<uint32_t>(w), wh = static_cast<uint32_t>(w >> 32);
      keep |= (wl >= thr ? 1u : 0u) << (4 * q);
      keep |= ((wl << 16) >= thr ? 1u : 0u) << (4 * q + 1);
      keep |= (wh >= thr ? 1u : 0u) << (4 * q + 2);
      keep |= ((wh << 16) >= thr ? 1u : 0u) << (4 * q + 3);
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < 32; ++j) keep |= (dropout_keep(seed, idx0 + j, thr) ? 1u : 0u) << j;
  }
  return keep;
}

// bias of 32 consecutive columns into registers (issued BEFORE the TMEM wait so the latency hides behind it);
// all lanes read the same addresses (one broadcast transaction per load)
__device__ __forceinline__ void epi_load_bias32(const EpiArgs& ep, int col0, int N, bool full, float (&b)[32]) {
  if (ep.bias == nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = 0.f;
  } else if (full) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + g * 4));
      b[g * 4] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = (col0 + j < N) ? __ldg(ep.bias + col0 + j) : 0.f;
  }
}

// Math of NG groups of 8 consecutive columns of one row: acc (TMEM registers) -> packed bf16 pairs (NG = 2: the
// 16-column steps of the 16-warp epilogue of the CTA-pair kernel, which keeps its live registers under the 112 a
// 576-thread CTA allows).  bias: the bias values (zeros when there is none); aux4: the thread's 16-byte pieces of the
// auxiliary operand, already in registers; keep: dropout keep bits, bit j = column j.
template <int MODE, int NG>
__device__ __forceinline__ void epi_compute_groups(const EpiArgs& ep, const uint32_t (&r)[8 * NG],
                                                   const float (&bias)[8 * NG], const uint4 (&aux4)[NG], uint32_t keep,
                                                   uint32_t (&outp)[4 * NG], uint32_t (&prep)[4 * NG]) {
  constexpr bool needs_aux = StagedEpi<MODE>::kAux;
  const float alpha = ep.alpha;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float v[8], a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(r[g * 8 + j]), alpha, bias[g * 8 + j]);
    if (needs_aux) {
      const float2 a0 = unpack_bf16x2(aux4[g].x), a1 = unpack_bf16x2(aux4[g].y), a2 = unpack_bf16x2(aux4[g].z),
                   a3 = unpack_bf16x2(aux4[g].w);
      a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y; a[6] = a3.x; a[7] = a3.y;
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = v[j];
      if (MODE == MTVAF_EPI_GELU) o[j] = gelu_fast(x);
      else if (MODE == MTVAF_EPI_TANH) o[j] = tanhf(x);
      else if (MODE == MTVAF_EPI_RESID)
        o[j] = ((keep >> (g * 8 + j)) & 1u) ? fmaf(x, ep.drop_scale, a[j]) : a[j];
      else if (MODE == MTVAF_EPI_GELU_GRAD) gelu_and_grad_fast(x, o[j], v[j]);   // v <- gelu'(x): the second output
      else if (MODE == MTVAF_EPI_MUL_DGELU) o[j] = x * dgelu_fast(a[j]);
      else if (MODE == MTVAF_EPI_MUL_DTANH) o[j] = x * (1.f - a[j] * a[j]);
      else if (MODE == MTVAF_EPI_MUL_AUX) o[j] = x * a[j];
      else o[j] = x;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) outp[g * 4 + j] = pack_bf16x2(o[2 * j], o[2 * j + 1]);
    if (StagedEpi<MODE>::kTwoOut) {
#pragma unroll
      for (int j = 0; j < 4; ++j) prep[g * 4 + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    }
  }
}

// bias of 16 consecutive columns into registers
__device__ __forceinline__ void epi_load_bias16(const EpiArgs& ep, int col0, int N, bool full, float (&b)[16]) {
  if (ep.bias == nullptr) {
#pragma unroll
    for (int j = 0; j < 16; ++j) b[j] = 0.f;
  } else if (full) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + g * 4));
      b[g * 4] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) b[j] = (col0 + j < N) ? __ldg(ep.bias + col0 + j) : 0.f;
  }
}

// byte offset of 16-byte piece `c16` (0..7) of row `row` (0..31) inside a SWIZZLE_128B [32][64 bf16] box
__device__ __forceinline__ uint32_t box_piece_off(int row, int c16) {
  return static_cast<uint32_t>(row * 128 + ((c16 ^ (row & 7)) << 4));
}

}  // namespace mtvaf
