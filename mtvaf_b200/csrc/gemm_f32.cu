// fp32 SIMT GEMM (FFMA) -- the PARITY-mode contraction (fp32 storage and math, so that logits/loss
// match the fp32 reference to 1e-4 through 12 layers, which tensor-core tf32/bf16 cannot) and the
// kernel for skinny heads (fc 768->11, projectors 6144->4).  Same operand-major conventions and fused
// epilogues as the tcgen05 kernel in gemm_tc.cu.
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile, double-buffered shared memory.
#include "common.cuh"
#include "../../include/mtvaf_b200.h"
#include "epilogue.cuh"

namespace mtvaf {

constexpr int SB_M = 128, SB_N = 128, SB_K = 16, S_THREADS = 256;

// loads a [rows=SB_M or SB_N][SB_K] operand tile into smem laid out [k][row] (row contiguous)
template <bool MN_MAJOR>
__device__ __forceinline__ void load_tile(const float* __restrict__ G, long long ld, int row0, int k0, int rows,
                                          int K, float (*S)[SB_M + 4], bool vec_ok) {
  const int tid = threadIdx.x;
  if (!MN_MAJOR) {
    // G[row][k], k contiguous: 128 rows x 16 k = 512 float4, 2 per thread
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * S_THREADS;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      const int gr = row0 + r, gk = k0 + kq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rows) {
        const float* p = G + (long long)gr * ld + gk;
        if (vec_ok && gk + 3 < K) {
          v = *reinterpret_cast<const float4*>(p);
        } else {
          if (gk + 0 < K) v.x = p[0];
          if (gk + 1 < K) v.y = p[1];
          if (gk + 2 < K) v.z = p[2];
          if (gk + 3 < K) v.w = p[3];
        }
      }
      S[kq + 0][r] = v.x; S[kq + 1][r] = v.y; S[kq + 2][r] = v.z; S[kq + 3][r] = v.w;
    }
  } else {
    // G[k][row], row contiguous: 16 k x 128 rows = 512 float4, 2 per thread
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * S_THREADS;
      const int k = idx >> 5, rq = (idx & 31) * 4;
      const int gk = k0 + k, gr = row0 + rq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gk < K) {
        const float* p = G + (long long)gk * ld + gr;
        if (vec_ok && gr + 3 < rows) {
          v = *reinterpret_cast<const float4*>(p);
        } else {
          if (gr + 0 < rows) v.x = p[0];
          if (gr + 1 < rows) v.y = p[1];
          if (gr + 2 < rows) v.z = p[2];
          if (gr + 3 < rows) v.w = p[3];
        }
      }
      *reinterpret_cast<float4*>(&S[k][rq]) = v;
    }
  }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(S_THREADS)
gemm_f32_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb, int M, int N,
                int K, int k_per_split, EpiArgs ep_in, int a_vec, int b_vec) {
  const EpiArgs ep = resolve_step(ep_in);
  __shared__ __align__(16) float As[2][SB_K][SB_M + 4];
  __shared__ __align__(16) float Bs[2][SB_K][SB_N + 4];
  const int m0 = blockIdx.y * SB_M, n0 = blockIdx.x * SB_N;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, each 8x8 outputs
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int buf = 0;
  load_tile<A_MN>(A, lda, m0, kbeg, M, kend, As[0], a_vec);
  load_tile<B_MN>(B, ldb, n0, kbeg, N, kend, Bs[0], b_vec);
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += SB_K) {
    if (k0 + SB_K < kend) {
      load_tile<A_MN>(A, lda, m0, k0 + SB_K, M, kend, As[buf ^ 1], a_vec);
      load_tile<B_MN>(B, ldb, n0, k0 + SB_K, N, kend, Bs[buf ^ 1], b_vec);
    }
#pragma unroll
    for (int k = 0; k < SB_K; ++k) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (col < N) epilogue_elem(ep, acc[i][j], row, col, N);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Skinny linear layer out[m, n] = sum_k x[m, k] w[n, k] + b[n] for N <= 48 (the tag head fc 768->11,
// bert_model.py:510, and the 12 gate projectors 6144->4 applied as one [48, 6144] matrix, :566-569).
// HBM/L2-bound: one warp per row, lanes split K in float4 pieces (coalesced 512 B per warp and step), the N
// accumulators live in registers and are reduced with shuffles at the end.  No split-K, no atomics:
// the result is bitwise reproducible, which the 128x128-tile kernel above cannot offer for these shapes
// without wasting > 90 % of its FMAs (N = 11) or serialising K on 8 blocks (M = 1024, K = 6144).
template <int NT>
__global__ void __launch_bounds__(256)
skinny_linear_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w, long long ldw,
                     const float* __restrict__ bias, int M, int N, int K, float* __restrict__ out, long long ldo) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.f;
    const float* xr = x + (long long)m * ldx;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 a = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (n < N) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(w + (long long)n * ldw + k));
          acc[n] = fmaf(a.x, b.x, acc[n]);
          acc[n] = fmaf(a.y, b.y, acc[n]);
          acc[n] = fmaf(a.z, b.z, acc[n]);
          acc[n] = fmaf(a.w, b.w, acc[n]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (n < N) {
        const float s = warp_sum(acc[n]);
        if (lane == 0) out[(long long)m * ldo + n] = s + (bias ? bias[n] : 0.f);
      }
    }
  }
}

// backward of the skinny layer, data gradient with the INPUT dropout of the head fused in (bert_model.py:506):
//   dx[m, k] = keep(seed, m*K + k) / (1-p) * sum_n dy[m, n] w[n, k]      written in the encoder's compute dtype
// warp per row, lanes own float4 pieces of K, W (N x K fp32, L1-resident) is streamed once per row.
template <typename TO, int NT>
__global__ void __launch_bounds__(256)
skinny_dgrad_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ w, long long ldw, int M,
                    int N, int K, TO* __restrict__ dx, long long lddx, const TO* __restrict__ tanh_out,
                    uint32_t drop_thr, float drop_scale, unsigned long long seed_in,
                    const unsigned long long* __restrict__ step) {
  const unsigned long long seed = drop_thr ? step_seed(seed_in, step) : seed_in;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    float g[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) g[n] = n < N ? dy[(long long)m * lddy + n] : 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (n < N) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(w + (long long)n * ldw + k));
          acc.x = fmaf(g[n], b.x, acc.x); acc.y = fmaf(g[n], b.y, acc.y);
          acc.z = fmaf(g[n], b.z, acc.z); acc.w = fmaf(g[n], b.w, acc.w);
        }
      }
      if (drop_thr) {
        const unsigned long long i0 = (unsigned long long)m * K + k;
        acc.x = dropout_keep(seed, i0 + 0, drop_thr) ? acc.x * drop_scale : 0.f;
        acc.y = dropout_keep(seed, i0 + 1, drop_thr) ? acc.y * drop_scale : 0.f;
        acc.z = dropout_keep(seed, i0 + 2, drop_thr) ? acc.z * drop_scale : 0.f;
        acc.w = dropout_keep(seed, i0 + 3, drop_thr) ? acc.w * drop_scale : 0.f;
      }
      if (tanh_out) {                                  // the layer's input was tanh(.): multiply by 1 - tanh^2
        const TO* t = tanh_out + (long long)m * lddx + k;
        const float t0 = to_f<TO>(t[0]), t1 = to_f<TO>(t[1]), t2 = to_f<TO>(t[2]), t3 = to_f<TO>(t[3]);
        acc.x *= 1.f - t0 * t0; acc.y *= 1.f - t1 * t1; acc.z *= 1.f - t2 * t2; acc.w *= 1.f - t3 * t3;
      }
      TO* o = dx + (long long)m * lddx + k;
      if constexpr (sizeof(TO) == 4) {
        *reinterpret_cast<float4*>(o) = acc;
      } else {
        uint2 u;
        u.x = pack_bf16x2(acc.x, acc.y);
        u.y = pack_bf16x2(acc.z, acc.w);
        *reinterpret_cast<uint2*>(o) = u;
      }
    }
  }
}

// weight gradient of the skinny layer: dw[n, k] += sum_m dy[m, n] x[m, k].  Each thread owns up to PPT
// (n, 8-wide k vector) pairs and walks the block's rows; x vectors are shared through L1 by the N threads of a
// k vector, dy entries are warp-broadcast loads.  One fp32 atomic per element and block at the end.
template <int PPT>
__global__ void __launch_bounds__(256)
skinny_wgrad_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx, int M,
                    int N, int K, float* __restrict__ dw, long long lddw) {
  const int kv = K >> 3;
  const int pairs = N * kv;
  float acc[PPT][8];
  int pn[PPT], pk[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int id = threadIdx.x + p * 256;
    pn[p] = id < pairs ? id / kv : -1;
    pk[p] = id < pairs ? (id - (id / kv) * kv) * 8 : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
  }
  const int rows_per = (M + gridDim.x - 1) / gridDim.x;
  const int m0 = blockIdx.x * rows_per, m1 = min(M, m0 + rows_per);
  // two rows per iteration: twice the loads in flight per thread (the loop is latency-bound, not bandwidth-bound)
  int m = m0;
  for (; m + 1 < m1; m += 2) {
    const float* xr0 = x + (long long)m * ldx;
    const float* xr1 = xr0 + ldx;
    const float* gr0 = dy + (long long)m * lddy;
    const float* gr1 = gr0 + lddy;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      if (pn[p] >= 0) {
        const float g0 = gr0[pn[p]], g1 = gr1[pn[p]];
        const float4 a0 = *reinterpret_cast<const float4*>(xr0 + pk[p]);
        const float4 b0 = *reinterpret_cast<const float4*>(xr0 + pk[p] + 4);
        const float4 a1 = *reinterpret_cast<const float4*>(xr1 + pk[p]);
        const float4 b1 = *reinterpret_cast<const float4*>(xr1 + pk[p] + 4);
        acc[p][0] = fmaf(g0, a0.x, acc[p][0]); acc[p][1] = fmaf(g0, a0.y, acc[p][1]);
        acc[p][2] = fmaf(g0, a0.z, acc[p][2]); acc[p][3] = fmaf(g0, a0.w, acc[p][3]);
        acc[p][4] = fmaf(g0, b0.x, acc[p][4]); acc[p][5] = fmaf(g0, b0.y, acc[p][5]);
        acc[p][6] = fmaf(g0, b0.z, acc[p][6]); acc[p][7] = fmaf(g0, b0.w, acc[p][7]);
        acc[p][0] = fmaf(g1, a1.x, acc[p][0]); acc[p][1] = fmaf(g1, a1.y, acc[p][1]);
        acc[p][2] = fmaf(g1, a1.z, acc[p][2]); acc[p][3] = fmaf(g1, a1.w, acc[p][3]);
        acc[p][4] = fmaf(g1, b1.x, acc[p][4]); acc[p][5] = fmaf(g1, b1.y, acc[p][5]);
        acc[p][6] = fmaf(g1, b1.z, acc[p][6]); acc[p][7] = fmaf(g1, b1.w, acc[p][7]);
      }
    }
  }
  for (; m < m1; ++m) {
    const float* xr = x + (long long)m * ldx;
    const float* gr = dy + (long long)m * lddy;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      if (pn[p] >= 0) {
        const float g = gr[pn[p]];
        const float4 a = *reinterpret_cast<const float4*>(xr + pk[p]);
        const float4 b = *reinterpret_cast<const float4*>(xr + pk[p] + 4);
        acc[p][0] = fmaf(g, a.x, acc[p][0]); acc[p][1] = fmaf(g, a.y, acc[p][1]);
        acc[p][2] = fmaf(g, a.z, acc[p][2]); acc[p][3] = fmaf(g, a.w, acc[p][3]);
        acc[p][4] = fmaf(g, b.x, acc[p][4]); acc[p][5] = fmaf(g, b.y, acc[p][5]);
        acc[p][6] = fmaf(g, b.z, acc[p][6]); acc[p][7] = fmaf(g, b.w, acc[p][7]);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    if (pn[p] >= 0) {
      float* o = dw + (long long)pn[p] * lddw + pk[p];
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(o + j, acc[p][j]);
    }
  }
}

}  // namespace mtvaf

using namespace mtvaf;

extern "C" int mtvaf_gemm_f32(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                              int b_mn_major, int M, int N, int K, const MtvafEpilogue* epi, int splits,
                              void* stream) {
  MTVAF_REQUIRE(A && B && epi, "mtvaf_gemm_f32: null argument");
  MTVAF_REQUIRE(M > 0 && N > 0 && K > 0, "mtvaf_gemm_f32: empty problem M=%d N=%d K=%d", M, N, K);
  if (splits < 1) splits = 1;
  int k_per = ((K + splits - 1) / splits + SB_K - 1) / SB_K * SB_K;
  splits = (K + k_per - 1) / k_per;
  EpiArgs ep;
  int rc = make_epi_args(*epi, MTVAF_F32, M, N, splits, &ep);
  if (rc) return rc;
  const float* a = static_cast<const float*>(A);
  const float* b = static_cast<const float*>(B);
  const int a_vec = (reinterpret_cast<uintptr_t>(a) % 16 == 0) && (lda % 4 == 0);
  const int b_vec = (reinterpret_cast<uintptr_t>(b) % 16 == 0) && (ldb % 4 == 0);
  dim3 grid((N + SB_N - 1) / SB_N, (M + SB_M - 1) / SB_M, splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!a_mn_major && !b_mn_major)
    gemm_f32_kernel<false, false><<<grid, S_THREADS, 0, st>>>(a, lda, b, ldb, M, N, K, k_per, ep, a_vec, b_vec);
  else if (!a_mn_major && b_mn_major)
    gemm_f32_kernel<false, true><<<grid, S_THREADS, 0, st>>>(a, lda, b, ldb, M, N, K, k_per, ep, a_vec, b_vec);
  else if (a_mn_major && b_mn_major)
    gemm_f32_kernel<true, true><<<grid, S_THREADS, 0, st>>>(a, lda, b, ldb, M, N, K, k_per, ep, a_vec, b_vec);
  else
    gemm_f32_kernel<true, false><<<grid, S_THREADS, 0, st>>>(a, lda, b, ldb, M, N, K, k_per, ep, a_vec, b_vec);
  MTVAF_LAUNCH_CHECK();
  if (epi->colsum) return mtvaf_colsum(epi->out, epi->ldo, epi->out_dtype, M, N, epi->colsum, stream);
  return 0;
}

extern "C" int mtvaf_skinny_linear_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                                       int M, int N, int K, float* out, int64_t ldo, void* stream) {
  MTVAF_REQUIRE(x && w && out && M > 0 && N > 0 && K > 0, "skinny_linear: bad argument");
  MTVAF_REQUIRE(N <= 48, "skinny_linear: N=%d > 48 (use mtvaf_gemm_f32)", N);
  MTVAF_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(w) % 16 == 0,
                "skinny_linear: K / leading dims must be multiples of 4 and x, w 16-byte aligned");
  long long blocks = ((long long)M + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N <= 16) skinny_linear_kernel<16><<<(int)blocks, 256, 0, st>>>(x, ldx, w, ldw, bias, M, N, K, out, ldo);
  else skinny_linear_kernel<48><<<(int)blocks, 256, 0, st>>>(x, ldx, w, ldw, bias, M, N, K, out, ldo);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_skinny_linear_dgrad(const float* dy, int64_t lddy, const float* w, int64_t ldw, int M, int N,
                                         int K, float p_drop, uint64_t seed, void* dx, int64_t lddx, int dx_dtype,
                                         const void* tanh_out, void* stream) {
  MTVAF_REQUIRE(dy && w && dx && M > 0 && N > 0 && K > 0, "skinny_linear_dgrad: bad argument");
  MTVAF_REQUIRE(N <= 16, "skinny_linear_dgrad: N=%d > 16 (use mtvaf_gemm_f32)", N);
  MTVAF_REQUIRE(K % 4 == 0 && ldw % 4 == 0 && lddx % 4 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(dx) % 16 == 0,
                "skinny_linear_dgrad: K / leading dims must be multiples of 4 and w, dx 16-byte aligned");
  MTVAF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "skinny_linear_dgrad: bad dropout p");
  uint32_t thr = 0;
  float scale = 1.f;
  if (p_drop > 0.f) {
    const double t = (double)p_drop * 4294967296.0;
    thr = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    scale = 1.f / (1.f - p_drop);
  }
  long long blocks = ((long long)M + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dx_dtype == MTVAF_BF16)
    skinny_dgrad_kernel<__nv_bfloat16, 16><<<(int)blocks, 256, 0, st>>>(dy, lddy, w, ldw, M, N, K, (__nv_bfloat16*)dx,
                                                                        lddx, (const __nv_bfloat16*)tanh_out, thr,
                                                                        scale, seed, step_source());
  else
    skinny_dgrad_kernel<float, 16><<<(int)blocks, 256, 0, st>>>(dy, lddy, w, ldw, M, N, K, (float*)dx, lddx,
                                                                (const float*)tanh_out, thr, scale, seed,
                                                                step_source());
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_skinny_linear_wgrad(const float* dy, int64_t lddy, const float* x, int64_t ldx, int M, int N,
                                         int K, float* dw, int64_t lddw, void* stream) {
  MTVAF_REQUIRE(dy && x && dw && M > 0 && N > 0 && K > 0, "skinny_linear_wgrad: bad argument");
  MTVAF_REQUIRE(K % 8 == 0 && ldx % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0,
                "skinny_linear_wgrad: K %% 8 == 0, ldx %% 4 == 0 and x 16-byte aligned required");
  MTVAF_REQUIRE((long long)N * (K / 8) <= 6 * 256, "skinny_linear_wgrad: N*K/8 = %lld > 1536 (use mtvaf_gemm_f32)",
                (long long)N * (K / 8));
  int blocks = sm_count() * 4;                  // 4 resident blocks per SM hide the load latency of the row walk
  if (blocks > (M + 7) / 8) blocks = (M + 7) / 8;
  if (blocks < 1) blocks = 1;
  skinny_wgrad_kernel<6><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, lddy, x, ldx, M, N, K, dw, lddw);
  MTVAF_LAUNCH_CHECK();
  return 0;
}
