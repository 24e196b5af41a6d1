// Heads of the span variant TVNetSAModel (models/bert_model.py:113-190, 323-376): distant cross-entropy on the
// start / end logits, ragged span gather + self-attentive pooling (forward and backward), mean cross-entropy of
// the polarity classifier.  All fp32 (the heads are tiny next to the encoder); gather indices are bit-exact.
//
// Span indexing follows get_span_representation (:147-172) literally: spans address the COMPACTED token stream
// (the tokens with attention_mask == 1, sentence after sentence), every index is clamped to the last token of the
// stream, and a span of width w contributes its first w gathered rows (rows j >= w carry the additive -10000
// of get_self_att_representation :174-181, i.e. exactly zero probability in fp32).  Attention masks are
// left-aligned (ones then zeros), as everywhere on this path (modules/dataset.py:398-418).
#include "common.cuh"
#include "../../include/mtvaf_b200.h"

namespace mtvaf {

// ------------------------------------------------------------------ sentence lengths / offsets of the stream
// ws[0..B) = lengths, ws[B..2B) = exclusive prefix sums, ws[2B] = total number of tokens in the stream
__global__ void span_offsets_kernel(const long long* __restrict__ mask, int B, int L, int* __restrict__ ws) {
  __shared__ int lens[1024];
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int n = 0;
    for (int l = 0; l < L; ++l) n += (mask[(long long)b * L + l] != 0);
    lens[b] = n;
    ws[b] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < B; ++b) { ws[B + b] = run; run += lens[b]; }
    ws[2 * B] = run;
  }
}

// row of the [B*L, H] activation that holds element g of the compacted stream
__device__ __forceinline__ int stream_row(const int* __restrict__ ws, int B, int L, int g) {
  const int* off = ws + B;
  int lo = 0, hi = B - 1;                             // last sentence whose offset is <= g
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= g) lo = mid; else hi = mid - 1;
  }
  // skip empty sentences that share the same offset: the owner is the one with g - off < len
  while (lo > 0 && g - off[lo] >= ws[lo]) --lo;
  return lo * L + (g - off[lo]);
}

constexpr int SP_MAX_VEC = 4;   // H <= 1024

template <int NV>
__device__ __forceinline__ void sp_load(const float* __restrict__ p, int H, int lane, float (&v)[NV][8]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < H) Vec8<float>::load(p + c, v[i]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
}
template <int NV>
__device__ __forceinline__ float sp_dot(const float (&a)[NV][8], const float (&b)[NV][8]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(a[i][j], b[i][j], s);
  return warp_sum(s);
}

// one warp per span: pooled[s, :] = sum_j softmax_j(row_j . w + b) row_j
template <int NV>
__global__ void __launch_bounds__(128)
span_pool_fwd_kernel(const float* __restrict__ seq, const int* __restrict__ ws, const long long* __restrict__ starts,
                     const long long* __restrict__ ends, const float* __restrict__ w_u,
                     const float* __restrict__ b_u, int B, int L, int M, int H, float* __restrict__ pooled) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (s >= B * M) return;
  const int n = s / M;
  const int total = ws[2 * B];
  const int g0 = (int)starts[s] + ws[B + n];
  const int width = (int)(ends[s] - starts[s]) + 1;
  float w[NV][8], acc[NV][8];
  sp_load<NV>(w_u, H, lane, w);
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float bias = b_u[0];
  float mx = -INFINITY, sum = 0.f;                    // online softmax over the span
  for (int j = 0; j < width; ++j) {
    const int g = min(g0 + j, total - 1);
    float x[NV][8];
    sp_load<NV>(seq + (long long)stream_row(ws, B, L, g) * H, H, lane, x);
    const float sc = sp_dot<NV>(x, w) + bias;
    const float nm = fmaxf(mx, sc);
    const float corr = __expf(mx - nm), p = __expf(sc - nm);
    sum = sum * corr + p;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(p, x[i][k], acc[i][k] * corr);
    mx = nm;
  }
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < H) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = acc[i][k] * inv;
      Vec8<float>::store(pooled + (long long)s * H + c, o);
    }
  }
}

// backward: d_row_j += p_j d_pooled + p_j (g_j - G) w ;  d_w += sum_j ds_j row_j ;  d_b += sum_j ds_j
// with g_j = d_pooled . row_j, G = sum_j p_j g_j, ds_j = p_j (g_j - G)
template <int NV>
__global__ void __launch_bounds__(128)
span_pool_bwd_kernel(const float* __restrict__ d_pooled, const float* __restrict__ seq, const int* __restrict__ ws,
                     const long long* __restrict__ starts, const long long* __restrict__ ends,
                     const float* __restrict__ w_u, const float* __restrict__ b_u, int B, int L, int M, int H,
                     float* __restrict__ d_seq, float* __restrict__ d_w, float* __restrict__ d_b) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (s >= B * M) return;
  const int n = s / M;
  const int total = ws[2 * B];
  const int g0 = (int)starts[s] + ws[B + n];
  const int width = (int)(ends[s] - starts[s]) + 1;
  float w[NV][8], dp[NV][8], dw[NV][8];
  sp_load<NV>(w_u, H, lane, w);
  sp_load<NV>(d_pooled + (long long)s * H, H, lane, dp);
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dw[i][j] = 0.f;
  const float bias = b_u[0];
  // pass 1: softmax statistics and G
  float mx = -INFINITY, sum = 0.f, G = 0.f;
  for (int j = 0; j < width; ++j) {
    const int g = min(g0 + j, total - 1);
    float x[NV][8];
    sp_load<NV>(seq + (long long)stream_row(ws, B, L, g) * H, H, lane, x);
    const float sc = sp_dot<NV>(x, w) + bias;
    const float gj = sp_dot<NV>(x, dp);
    const float nm = fmaxf(mx, sc);
    const float corr = __expf(mx - nm), p = __expf(sc - nm);
    sum = sum * corr + p;
    G = G * corr + p * gj;
    mx = nm;
  }
  if (!(sum > 0.f)) return;
  G /= sum;
  // pass 2: gradients
  float db = 0.f;
  for (int j = 0; j < width; ++j) {
    const int g = min(g0 + j, total - 1);
    const long long r = stream_row(ws, B, L, g);
    float x[NV][8];
    sp_load<NV>(seq + r * H, H, lane, x);
    const float sc = sp_dot<NV>(x, w) + bias;
    const float gj = sp_dot<NV>(x, dp);
    const float p = __expf(sc - mx) / sum;
    const float ds = p * (gj - G);
    db += ds;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      if (c < H) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          dw[i][k] = fmaf(ds, x[i][k], dw[i][k]);
          o[k] = fmaf(p, dp[i][k], ds * w[i][k]);
        }
        float* dst = d_seq + r * H + c;               // 16-byte vector reductions
        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
        atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(o[4], o[5], o[6], o[7]));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < H) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(d_w + c + k, dw[i][k]);
    }
  }
  if (lane == 0) atomicAdd(d_b, db);
}

// ------------------------------------------------------------------ distant cross-entropy (:183-192, no mask)
// logits / dlogits are addressed with an element stride (start and end logits are the two columns of one [T, 2]
// matrix); loss[0] += scale * mean_b( -sum_l pos log_softmax(x)_l / sum_l pos );  dlogits = d(scale * that)/dx
__global__ void __launch_bounds__(256)
distant_ce_kernel(const float* __restrict__ logits, long long stride, const long long* __restrict__ pos, int B, int L,
                  float scale, float* __restrict__ loss, float* __restrict__ dlogits) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* x = logits + (long long)b * L * stride;
  const long long* p = pos + (long long)b * L;
  float mx = -INFINITY;
  for (int l = threadIdx.x; l < L; l += blockDim.x) mx = fmaxf(mx, x[l * stride]);
  mx = block_max(mx, red);
  float se = 0.f, np = 0.f, sp = 0.f;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float v = x[l * stride];
    se += __expf(v - mx);
    const float q = (float)p[l];
    np += q;
    sp += q * v;
  }
  se = block_sum(se, red);
  np = block_sum(np, red);
  sp = block_sum(sp, red);
  const float lse = mx + __logf(se);
  // -sum pos (x - lse) / np = (np * lse - sp) / np
  if (threadIdx.x == 0) atomicAdd(loss, scale * (np * lse - sp) / np / (float)B);
  if (dlogits) {
    float* d = dlogits + (long long)b * L * stride;
    const float f = scale / (float)B;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const float sm = __expf(x[l * stride] - lse);
      d[l * stride] = f * (sm - (float)p[l] / np);
    }
  }
}

// ------------------------------------------------------------------ mean cross-entropy over N rows of C classes
__global__ void __launch_bounds__(256)
ce_mean_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int N, int C, float scale,
               float* __restrict__ loss, float* __restrict__ dlogits) {
  __shared__ float red[32];
  float part = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
    const float* x = logits + (long long)r * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, x[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += __expf(x[c] - mx);
    const float lse = mx + __logf(se);
    const int y = (int)labels[r];
    part += lse - x[y];
    if (dlogits) {
      for (int c = 0; c < C; ++c)
        dlogits[(long long)r * C + c] = scale / (float)N * (__expf(x[c] - lse) - (c == y ? 1.f : 0.f));
    }
  }
  part = block_sum(part, red);
  if (threadIdx.x == 0) atomicAdd(loss, scale * part / (float)N);
}

}  // namespace mtvaf

using namespace mtvaf;

extern "C" int mtvaf_span_offsets(const int64_t* attention_mask, int B, int L, int32_t* workspace, void* stream) {
  MTVAF_REQUIRE(attention_mask && workspace && B > 0 && B <= 1024 && L > 0, "span_offsets: bad argument (B <= 1024)");
  span_offsets_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const long long*)attention_mask, B, L, workspace);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_span_pool_fwd(const float* seq, const int32_t* workspace, const int64_t* span_starts,
                                   const int64_t* span_ends, const float* w_unary, const float* b_unary, int B, int L,
                                   int M, int H, float* pooled, void* stream) {
  MTVAF_REQUIRE(seq && workspace && span_starts && span_ends && w_unary && b_unary && pooled, "span_pool_fwd: null argument");
  MTVAF_REQUIRE(B > 0 && L > 0 && M > 0 && H % 8 == 0 && H <= SP_MAX_VEC * 256, "span_pool_fwd: bad shape");
  const int grid = (B * M + 3) / 4;
  const int nv = (H + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
#define MTVAF_SPF(NV_)                                                                                         \
  span_pool_fwd_kernel<NV_><<<grid, 128, 0, st>>>(seq, workspace, (const long long*)span_starts,               \
                                                  (const long long*)span_ends, w_unary, b_unary, B, L, M, H, pooled)
  if (nv == 1) MTVAF_SPF(1); else if (nv == 2) MTVAF_SPF(2); else if (nv == 3) MTVAF_SPF(3); else MTVAF_SPF(4);
#undef MTVAF_SPF
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_span_pool_bwd(const float* d_pooled, const float* seq, const int32_t* workspace,
                                   const int64_t* span_starts, const int64_t* span_ends, const float* w_unary,
                                   const float* b_unary, int B, int L, int M, int H, float* d_seq, float* d_w_unary,
                                   float* d_b_unary, void* stream) {
  MTVAF_REQUIRE(d_pooled && seq && workspace && span_starts && span_ends && w_unary && b_unary && d_seq && d_w_unary &&
                    d_b_unary, "span_pool_bwd: null argument");
  MTVAF_REQUIRE(B > 0 && L > 0 && M > 0 && H % 8 == 0 && H <= SP_MAX_VEC * 256, "span_pool_bwd: bad shape");
  const int grid = (B * M + 3) / 4;
  const int nv = (H + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
#define MTVAF_SPB(NV_)                                                                                          \
  span_pool_bwd_kernel<NV_><<<grid, 128, 0, st>>>(d_pooled, seq, workspace, (const long long*)span_starts,      \
                                                  (const long long*)span_ends, w_unary, b_unary, B, L, M, H,   \
                                                  d_seq, d_w_unary, d_b_unary)
  if (nv == 1) MTVAF_SPB(1); else if (nv == 2) MTVAF_SPB(2); else if (nv == 3) MTVAF_SPB(3); else MTVAF_SPB(4);
#undef MTVAF_SPB
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_distant_ce_fwd_bwd(const float* logits, int64_t stride, const int64_t* positions, int B, int L,
                                        float scale, float* loss, float* dlogits, void* stream) {
  MTVAF_REQUIRE(logits && positions && loss && B > 0 && L > 0 && stride > 0, "distant_ce: bad argument");
  distant_ce_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(logits, stride, (const long long*)positions, B, L, scale, loss,
                                                         dlogits);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_ce_mean_fwd_bwd(const float* logits, const int64_t* labels, int N, int C, float scale,
                                     float* loss, float* dlogits, void* stream) {
  MTVAF_REQUIRE(logits && labels && loss && N > 0 && C > 0, "ce_mean: bad argument");
  int blocks = (N + 255) / 256;
  if (blocks > 64) blocks = 64;
  ce_mean_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, (const long long*)labels, N, C, scale, loss, dlogits);
  MTVAF_LAUNCH_CHECK();
  return 0;
}
