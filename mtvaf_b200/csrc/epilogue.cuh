// Fused GEMM epilogues shared by the tcgen05 (bf16) and SIMT (fp32) GEMM kernels.
// Each mode replaces the elementwise ATen kernels that follow a Linear on the reference path
// (bias add, GELU modeling_roberta.py:366, tanh bert_model.py:448, dropout+residual :297-298,
// probe squared norm probes/probe.py:76-78) or their backward.
#pragma once
#include "common.cuh"
#include "../../include/mtvaf_b200.h"

namespace mtvaf {

struct EpiArgs {
  int mode;
  int out_bf16;        // 1: out/out2 are bf16, 0: fp32
  int aux_bf16;        // dtype of aux
  int vec_ok;          // all pointers / leading dims allow 8-wide vector access
  int staged;          // bf16 out/out2/aux, 16-byte aligned, ld % 8 == 0: eligible for the TMA-staged epilogue
  void* out;
  long long ldo;
  const float* bias;
  const void* aux;
  long long ld_aux;
  void* out2;
  long long ld_out2;
  float* rowvec;
  float alpha;
  uint32_t drop_threshold;   // 0 = no dropout
  float drop_scale;          // 1/(1-p)
  unsigned long long seed;
  const unsigned long long* step;   // device step counter mixed into the seed (or NULL)
  float* colsum;                    // optional [N]: += column sums of `out` (see MtvafEpilogue)
};

inline int make_epi_args(const MtvafEpilogue& e, int operand_dtype, int M, int N, int splits, EpiArgs* o) {
  o->mode = e.mode;
  o->out_bf16 = (e.out_dtype == MTVAF_BF16);
  o->aux_bf16 = (operand_dtype == MTVAF_BF16);
  o->out = e.out; o->ldo = e.ldo; o->bias = e.bias; o->aux = e.aux; o->ld_aux = e.ld_aux;
  o->out2 = e.out2; o->ld_out2 = e.ld_out2; o->rowvec = e.rowvec;
  o->alpha = (e.alpha == 0.f) ? 1.f : e.alpha;
  o->seed = e.seed;
  o->colsum = e.colsum;
  o->step = step_source();
  o->drop_threshold = 0; o->drop_scale = 1.f;
  MTVAF_REQUIRE(e.mode >= 0 && e.mode <= MTVAF_EPI_MUL_AUX, "bad epilogue mode %d", e.mode);
  if (e.mode == MTVAF_EPI_ATOMIC_F32) o->out_bf16 = 0;
  if (e.mode != MTVAF_EPI_SQNORM) MTVAF_REQUIRE(e.out != nullptr, "epilogue: out is NULL");
  if (e.mode == MTVAF_EPI_RESID || e.mode == MTVAF_EPI_MUL_DGELU || e.mode == MTVAF_EPI_MUL_DTANH ||
      e.mode == MTVAF_EPI_MUL_AUX)
    MTVAF_REQUIRE(e.aux != nullptr, "epilogue mode %d needs aux", e.mode);
  if (e.mode == MTVAF_EPI_GELU_GRAD) MTVAF_REQUIRE(e.out2 != nullptr, "epilogue GELU_GRAD needs out2");
  if (e.mode == MTVAF_EPI_SQNORM || e.mode == MTVAF_EPI_ROWSCALE)
    MTVAF_REQUIRE(e.rowvec != nullptr, "epilogue mode %d needs rowvec", e.mode);
  if (e.colsum)
    MTVAF_REQUIRE(e.mode != MTVAF_EPI_ATOMIC_F32 && e.mode != MTVAF_EPI_SQNORM && e.out != nullptr,
                  "epilogue: colsum needs a stored output");
  if (splits > 1)
    MTVAF_REQUIRE(e.mode == MTVAF_EPI_ATOMIC_F32 || (e.mode == MTVAF_EPI_SQNORM && e.out == nullptr),
                  "split-K needs an accumulating epilogue");
  if (e.mode == MTVAF_EPI_RESID && e.p_drop > 0.f) {
    MTVAF_REQUIRE(e.p_drop < 1.f, "dropout p must be < 1");
    double t = (double)e.p_drop * 4294967296.0;
    o->drop_threshold = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    o->drop_scale = 1.f / (1.f - e.p_drop);
  }
  auto al = [](const void* p, long long ld, int bf16) {
    if (!p) return true;
    const int w = bf16 ? 8 : 4;
    return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % w == 0);
  };
  o->vec_ok = al(o->out, o->ldo, o->out_bf16) && al(o->aux, o->ld_aux, o->aux_bf16) &&
              al(o->out2, o->ld_out2, o->out_bf16) && (!o->bias || reinterpret_cast<uintptr_t>(o->bias) % 16 == 0);
  o->staged = o->vec_ok && o->out_bf16 && o->out != nullptr && (!o->aux || o->aux_bf16) && splits <= 1;
  return 0;
}

// kernels call this once at entry: folds the device step counter into the by-value seed (one global load per
// thread instead of one per element)
__device__ __forceinline__ EpiArgs resolve_step(const EpiArgs& in) {
  EpiArgs e = in;
  if (in.drop_threshold) e.seed = step_seed(in.seed, in.step);
  e.step = nullptr;
  return e;
}

__device__ __forceinline__ float epi_load(const void* p, int bf16, long long idx) {
  return bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx])
              : reinterpret_cast<const float*>(p)[idx];
}
__device__ __forceinline__ void epi_store(void* p, int bf16, long long idx, float v) {
  if (bf16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[idx] = v;
}
__device__ __forceinline__ void epi_load8(const void* p, int bf16, long long idx, float (&v)[8]) {
  if (bf16) Vec8<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(p) + idx, v);
  else Vec8<float>::load(reinterpret_cast<const float*>(p) + idx, v);
}
__device__ __forceinline__ void epi_store8(void* p, int bf16, long long idx, const float (&v)[8]) {
  if (bf16) Vec8<__nv_bfloat16>::store(reinterpret_cast<__nv_bfloat16*>(p) + idx, v);
  else Vec8<float>::store(reinterpret_cast<float*>(p) + idx, v);
}

// scalar math of one output element; `a` = aux value (if the mode uses one).
// MODE >= 0 fixes the epilogue at compile time (lean code for the hot kernels); MODE < 0 reads ep.mode.
template <int MODE = -1, bool FAST = false>
__device__ __forceinline__ float epi_math(const EpiArgs& ep, float v, float a, int row, int col, int N,
                                          float& pre_out) {
  const int mode = (MODE >= 0) ? MODE : ep.mode;
  switch (mode) {
    case MTVAF_EPI_GELU: pre_out = v; return FAST ? gelu_fast(v) : gelu_erf(v);
    case MTVAF_EPI_GELU_GRAD: pre_out = FAST ? dgelu_fast(v) : dgelu_erf(v); return FAST ? gelu_fast(v) : gelu_erf(v);
    case MTVAF_EPI_MUL_AUX: return v * a;
    case MTVAF_EPI_TANH: return tanhf(v);
    case MTVAF_EPI_RESID:
      if (ep.drop_threshold) {
        const bool keep = dropout_keep(ep.seed, (unsigned long long)row * (unsigned long long)N + col,
                                       ep.drop_threshold);
        v = keep ? v * ep.drop_scale : 0.f;
      }
      return v + a;
    case MTVAF_EPI_MUL_DGELU: return v * (FAST ? dgelu_fast(a) : dgelu_erf(a));
    case MTVAF_EPI_MUL_DTANH: return v * (1.f - a * a);
    default: return v;
  }
}

// One thread owns 32 consecutive columns [col0, col0+32) of output row `row` (tcgen05 epilogue).
template <int MODE = -1>
__device__ __forceinline__ void epilogue_row32(const EpiArgs& ep, const uint32_t (&r)[32], int row, int col0, int M,
                                               int N, float& rowacc) {
  if (row >= M) return;
  const int mode = (MODE >= 0) ? MODE : ep.mode;
  const bool needs_aux = (mode == MTVAF_EPI_RESID || mode == MTVAF_EPI_MUL_DGELU || mode == MTVAF_EPI_MUL_DTANH ||
                          mode == MTVAF_EPI_MUL_AUX);
  const float rs = (mode == MTVAF_EPI_ROWSCALE) ? ep.rowvec[row] : 1.f;
  const bool full = (col0 + 32 <= N) && ep.vec_ok;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = col0 + g * 8;
    if (c >= N) break;
    float v[8], a[8], pre[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) * ep.alpha * rs;
    if (full) {
      if (ep.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + c + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (needs_aux) epi_load8(ep.aux, ep.aux_bf16, (long long)row * ep.ld_aux + c, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        pre[j] = v[j];
        v[j] = epi_math<MODE>(ep, v[j], needs_aux ? a[j] : 0.f, row, c + j, N, pre[j]);
      }
      if (mode == MTVAF_EPI_ATOMIC_F32) {
        // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the scalar atomics' instruction count
        float* o = reinterpret_cast<float*>(ep.out) + (long long)row * ep.ldo + c;
        atomicAdd(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
        atomicAdd(reinterpret_cast<float4*>(o + 4), make_float4(v[4], v[5], v[6], v[7]));
      } else {
        if (mode == MTVAF_EPI_SQNORM) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rowacc += v[j] * v[j];
        }
        if (ep.out) epi_store8(ep.out, ep.out_bf16, (long long)row * ep.ldo + c, v);
        if ((mode == MTVAF_EPI_GELU || mode == MTVAF_EPI_GELU_GRAD) && ep.out2)
          epi_store8(ep.out2, ep.out_bf16, (long long)row * ep.ld_out2 + c, pre);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cc = c + j;
        if (cc >= N) break;
        float x = v[j] + (ep.bias ? ep.bias[cc] : 0.f);
        float p = x;
        const float av = needs_aux ? epi_load(ep.aux, ep.aux_bf16, (long long)row * ep.ld_aux + cc) : 0.f;
        x = epi_math<MODE>(ep, x, av, row, cc, N, p);
        if (mode == MTVAF_EPI_ATOMIC_F32) {
          atomicAdd(reinterpret_cast<float*>(ep.out) + (long long)row * ep.ldo + cc, x);
        } else {
          if (mode == MTVAF_EPI_SQNORM) rowacc += x * x;
          if (ep.out) epi_store(ep.out, ep.out_bf16, (long long)row * ep.ldo + cc, x);
          if ((mode == MTVAF_EPI_GELU || mode == MTVAF_EPI_GELU_GRAD) && ep.out2)
            epi_store(ep.out2, ep.out_bf16, (long long)row * ep.ld_out2 + cc, p);
        }
      }
    }
  }
}

template <int MODE = -1>
__device__ __forceinline__ void epilogue_row_finish(const EpiArgs& ep, int row, int M, float rowacc) {
  const int mode = (MODE >= 0) ? MODE : ep.mode;
  if (mode == MTVAF_EPI_SQNORM && row < M) atomicAdd(ep.rowvec + row, rowacc);
}

// element-wise form used by the SIMT fp32 GEMM
__device__ __forceinline__ void epilogue_elem(const EpiArgs& ep, float acc, int row, int col, int N) {
  const bool needs_aux = (ep.mode == MTVAF_EPI_RESID || ep.mode == MTVAF_EPI_MUL_DGELU ||
                          ep.mode == MTVAF_EPI_MUL_DTANH || ep.mode == MTVAF_EPI_MUL_AUX);
  float x = acc * ep.alpha;
  if (ep.mode == MTVAF_EPI_ROWSCALE) x *= ep.rowvec[row];
  if (ep.bias) x += ep.bias[col];
  float p = x;
  const float av = needs_aux ? epi_load(ep.aux, ep.aux_bf16, (long long)row * ep.ld_aux + col) : 0.f;
  x = epi_math(ep, x, av, row, col, N, p);
  if (ep.mode == MTVAF_EPI_ATOMIC_F32) {
    atomicAdd(reinterpret_cast<float*>(ep.out) + (long long)row * ep.ldo + col, x);
    return;
  }
  if (ep.mode == MTVAF_EPI_SQNORM) atomicAdd(ep.rowvec + row, x * x);
  if (ep.out) epi_store(ep.out, ep.out_bf16, (long long)row * ep.ldo + col, x);
  if ((ep.mode == MTVAF_EPI_GELU || ep.mode == MTVAF_EPI_GELU_GRAD) && ep.out2)
    epi_store(ep.out2, ep.out_bf16, (long long)row * ep.ld_out2 + col, p);
}

}  // namespace mtvaf
