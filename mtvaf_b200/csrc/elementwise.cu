// Memory-bound helpers: dtype casts, column sums (bias gradients), 4-way means of the visual prompt,
// MSE (probe loss), loss combination without host sync, fused AdamW.  All are HBM-bound: 16-byte
// vector loads, grid-stride loops sized to the SM count, warp-shuffle / shared-memory reductions.
#include "common.cuh"
#include "../../include/mtvaf_b200.h"

namespace mtvaf {

// ------------------------------------------------------------------ casts
__global__ void cast_f32_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, long long n) {
  const long long n8 = n / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    Vec8<float>::load(s + i * 8, v);
    Vec8<__nv_bfloat16>::store(d + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < n - n8 * 8) d[n8 * 8 + threadIdx.x] = __float2bfloat16_rn(s[n8 * 8 + threadIdx.x]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ s, float* __restrict__ d, long long n) {
  const long long n8 = n / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    Vec8<__nv_bfloat16>::load(s + i * 8, v);
    Vec8<float>::store(d + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < n - n8 * 8) d[n8 * 8 + threadIdx.x] = __bfloat162float(s[n8 * 8 + threadIdx.x]);
}

static int grid_for(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  long long cap = (long long)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------ column sum: db[n] += sum_m dy[m][n]
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ dy, long long ld, int M, int N, int rows_per_block,
                              float* __restrict__ db) {
  __shared__ float red[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const int c = blockIdx.x * 64 + tx * 2;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (c < N) {
    const bool two = (c + 1 < N);
    for (int r = r0 + ty; r < r1; r += 8) {
      const T* p = dy + (long long)r * ld + c;
      s0 += to_f<T>(p[0]);
      if (two) s1 += to_f<T>(p[1]);
    }
  }
  red[ty][tx * 2] = s0;
  red[ty][tx * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < N) atomicAdd(db + cc, s);
  }
}

// vector form: a warp covers 256 columns of one row with 16-byte (bf16) / 32-byte (fp32) loads, 8 warps take
// interleaved rows, 4 rows in flight per thread; used whenever N % 8 == 0 and the rows are 16-byte aligned
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ dy, long long ld, int M, int N, int rows_per_block, float* __restrict__ db) {
  __shared__ float red[8][256 + 8];
  using V = Vec8<T>;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + tx * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < N) {
    int r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      typename V::Raw v0 = V::load_raw(dy + (long long)r * ld + c);
      typename V::Raw v1 = V::load_raw(dy + (long long)(r + 8) * ld + c);
      typename V::Raw v2 = V::load_raw(dy + (long long)(r + 16) * ld + c);
      typename V::Raw v3 = V::load_raw(dy + (long long)(r + 24) * ld + c);
      float f0[8], f1[8], f2[8], f3[8];
      V::cvt(v0, f0); V::cvt(v1, f1); V::cvt(v2, f2); V::cvt(v3, f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
    }
    for (; r < r1; r += 8) {
      float f0[8];
      V::load(dy + (long long)r * ld + c, f0);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f0[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    const int cc = blockIdx.x * 256 + threadIdx.x;
    if (cc < N) atomicAdd(db + cc, s);
  }
}

// ------------------------------------------------------------------ 4-way means of the prompt
// x: [rows, 4, W].  mode 0: y[row, w] = mean_r x[row, r, w]                 (bert_model.py:550)
//                   mode 1: y[row, r*S + c] = mean_i x[row, r, i*S + c], S = W/4   (bert_model.py:567)
template <typename T>
__global__ void mean4_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long rows, int W, int mode) {
  const long long total = rows * W;
  const int S = W / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long row = idx / W;
    const int w = (int)(idx - row * W);
    const T* base = x + row * 4 * W;
    float s;
    if (mode == 0) {
      s = to_f<T>(base[w]) + to_f<T>(base[W + w]) + to_f<T>(base[2 * W + w]) + to_f<T>(base[3 * W + w]);
    } else {
      const int r = w / S, c = w - r * S;
      const T* p = base + r * W + c;
      s = to_f<T>(p[0]) + to_f<T>(p[S]) + to_f<T>(p[2 * S]) + to_f<T>(p[3 * S]);
    }
    y[idx] = from_f<T>(s * 0.25f);
  }
}
// dx (fp32, [rows,4,W]) += broadcast(dy)/4
__global__ void mean4_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long rows, int W,
                                 int mode) {
  const long long total = rows * 4 * W;
  const int S = W / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long row = idx / (4 * W);
    const int rem = (int)(idx - row * 4 * W);
    const int r = rem / W, w = rem - r * W;
    float g;
    if (mode == 0) g = dy[row * W + w];
    else g = dy[row * W + r * S + (w % S)];
    dx[idx] += 0.25f * g;
  }
}

// ------------------------------------------------------------------ dropout apply / accumulate
template <typename T>
__global__ void dropout_apply_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, uint32_t thr,
                                     float scale, unsigned long long seed_in,
                                     const unsigned long long* __restrict__ step) {
  const unsigned long long seed = step_seed(seed_in, step);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = dropout_keep(seed, (unsigned long long)i, thr) ? from_f<T>(to_f<T>(x[i]) * scale) : from_f<T>(0.f);
}
// 8 elements per thread and step: 16-byte (bf16) / 2 x 16-byte (fp32) accesses, two mask hashes per vector
template <typename T>
__global__ void dropout_apply_vec_kernel(const T* __restrict__ x, T* __restrict__ y, long long n8, uint32_t thr,
                                         float scale, unsigned long long seed_in,
                                         const unsigned long long* __restrict__ step) {
  const unsigned long long seed = step_seed(seed_in, step);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    Vec8<T>::load(x + i * 8, v);
    const uint32_t keep = dropout_keep8(seed, (unsigned long long)i * 8, thr);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * scale : 0.f;
    Vec8<T>::store(y + i * 8, v);
  }
}
template <typename TD, typename TS>
__global__ void add_inplace_kernel(TD* __restrict__ dst, const TS* __restrict__ src, long long n, float alpha) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = from_f<TD>(to_f<TD>(dst[i]) + alpha * to_f<TS>(src[i]));
}

// ------------------------------------------------------------------ row scale / device-scalar scale
// y[m, n] = x[m, n] * alpha * rowscale[m]   (probe backward: dT = 2 g_m T_m)
template <typename T>
__global__ void rowscale_kernel(const T* __restrict__ x, const float* __restrict__ rs, T* __restrict__ y, long long M,
                                int N, float alpha) {
  const long long total = M * N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    y[i] = from_f<T>(to_f<T>(x[i]) * alpha * rs[i / N]);
}
// x[i] *= s[0] with s on the device (scaling head gradients by d(loss) without a host sync)
__global__ void scale_dev_kernel(float* __restrict__ x, long long n, const float* __restrict__ s) {
  const float f = s[0];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= f;
}

// ------------------------------------------------------------------ MSE (probe loss)
__global__ void mse_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float inv_n,
                           float* __restrict__ loss, float* __restrict__ da) {
  __shared__ float red[32];
  float s = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = a[i] - b[i];
    s += d * d;
    if (da) da[i] = 2.f * d * inv_n;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(loss, s * inv_n);
}

// ------------------------------------------------------------------ loss combination (device-side predicate)
__global__ void combine_loss_kernel(const float* crf_nll_sum, float inv_b, const float* prob_loss, float probe_coef,
                                    const float* img_losses, int n_img, float alpha, float* out, int* flag) {
  float l = crf_nll_sum[0] * inv_b;
  int f = 0;
  if (prob_loss) {
    const float pl = prob_loss[0];
    if (pl > 0.1f) { l += pl * probe_coef; f = 1; }     // probes/loss.py:14-16
  }
  float img = 0.f;
  for (int i = 0; i < n_img; ++i) img += img_losses[i];
  l += alpha * img;                                      // bert_model.py:525
  out[0] = l;
  if (flag) flag[0] = f;
}

// ------------------------------------------------------------------ AdamW (torch.optim.AdamW semantics)
// device-resident optimizer clock (CUDA-graph replay: nothing that changes per step may be a kernel argument)
struct AdamDyn { unsigned long long t; float lr_scale; float bc1; float bc2_sqrt; };
// t += 1; linear warm-up / linear decay factor of get_linear_schedule_with_warmup evaluated for the step about to
// be taken (modules/train.py:118-120,919-921) and the two Adam bias corrections
__global__ void adam_dyn_advance_kernel(AdamDyn* d, float b1, float b2, int warmup, int total) {
  const unsigned long long s = d->t;              // steps taken so far
  const unsigned long long t = s + 1;
  float scale = 1.f;
  if (total > 0) {
    if ((long long)s < warmup) scale = (float)s / fmaxf(1.f, (float)warmup);
    else scale = fmaxf(0.f, (float)((long long)total - (long long)s) / fmaxf(1.f, (float)(total - warmup)));
  }
  d->t = t;
  d->lr_scale = scale;
  d->bc1 = 1.f - powf(b1, (float)t);
  d->bc2_sqrt = sqrtf(1.f - powf(b2, (float)t));
}
__global__ void advance_step_kernel(unsigned long long* p) { *p += 1ull; }

__device__ __forceinline__ void adamw_one(float& w, float g, float& m, float& v, float lr, float b1, float b2,
                                          float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  const float grad = g * gscale;
  w *= (1.f - lr * wd);
  m = b1 * m + (1.f - b1) * grad;
  v = b2 * v + (1.f - b2) * grad * grad;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  w -= (lr / bc1) * (m / denom);
}
// 16-byte vector form (n4 = n / 4 quads; the caller guarantees 16-byte aligned bases); optionally refreshes the
// bf16 weight shadow and clears the gradient in the same pass (zero_grad), so a training step needs no
// separate cast or memset over the 125 M parameters
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
             float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale,
             __nv_bfloat16* __restrict__ bf, int zero_grad, const AdamDyn* __restrict__ dyn) {
  if (dyn) { lr *= dyn->lr_scale; bc1 = dyn->bc1; bc2_sqrt = dyn->bc2_sqrt; }
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 w4 = reinterpret_cast<float4*>(p)[i];
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i];
    float4 v4 = reinterpret_cast<float4*>(v)[i];
    adamw_one(w4.x, g4.x, m4.x, v4.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    adamw_one(w4.y, g4.y, m4.y, v4.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    adamw_one(w4.z, g4.z, m4.z, v4.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    adamw_one(w4.w, g4.w, m4.w, v4.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    reinterpret_cast<float4*>(p)[i] = w4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bf) {
      uint2 u;
      u.x = pack_bf16x2(w4.x, w4.y);
      u.y = pack_bf16x2(w4.z, w4.w);
      reinterpret_cast<uint2*>(bf)[i] = u;
    }
  }
  // scalar tail (n % 4 elements)
  const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    float w = p[t], mi = m[t], vi = v[t];
    adamw_one(w, g[t], mi, vi, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    p[t] = w; m[t] = mi; v[t] = vi;
    if (zero_grad) g[t] = 0.f;
    if (bf) bf[t] = __float2bfloat16_rn(w);
  }
}
__global__ void adamw_scalar_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                    float wd, float bc1, float bc2_sqrt, float gscale, __nv_bfloat16* __restrict__ bf,
                                    int zero_grad, const AdamDyn* __restrict__ dyn) {
  if (dyn) { lr *= dyn->lr_scale; bc1 = dyn->bc1; bc2_sqrt = dyn->bc2_sqrt; }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float w = p[i], mi = m[i], vi = v[i];
    adamw_one(w, g[i], mi, vi, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    p[i] = w; m[i] = mi; v[i] = vi;
    if (zero_grad) g[i] = 0.f;
    if (bf) bf[i] = __float2bfloat16_rn(w);
  }
}

}  // namespace mtvaf

using namespace mtvaf;

extern "C" int mtvaf_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n <= 0) return 0;
  MTVAF_REQUIRE(src && dst, "cast: null pointer");
  MTVAF_REQUIRE(reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0,
                "cast: pointers must be 16-byte aligned");
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  MTVAF_LAUNCH_CHECK();
  return 0;
}
extern "C" int mtvaf_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
  if (n <= 0) return 0;
  MTVAF_REQUIRE(src && dst, "cast: null pointer");
  MTVAF_REQUIRE(reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0,
                "cast: pointers must be 16-byte aligned");
  cast_bf16_f32_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, n);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_dropout_apply(const void* x, void* y, int64_t n, int dtype, float p_drop, uint64_t seed,
                                   void* stream) {
  if (n <= 0) return 0;
  MTVAF_REQUIRE(x && y && p_drop >= 0.f && p_drop < 1.f, "dropout_apply: bad argument");
  double t = (double)p_drop * 4294967296.0;
  const uint32_t thr = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
  const float scale = 1.f / (1.f - p_drop);
  const bool vec = n % 8 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0;
  if (vec && dtype == MTVAF_BF16)
    dropout_apply_vec_kernel<__nv_bfloat16><<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n / 8, thr, scale, seed, step_source());
  else if (vec)
    dropout_apply_vec_kernel<float><<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)x, (float*)y, n / 8, thr, scale, seed, step_source());
  else if (dtype == MTVAF_BF16)
    dropout_apply_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, thr, scale, seed, step_source());
  else
    dropout_apply_kernel<float><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, n, thr,
                                                                                    scale, seed, step_source());
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_add_inplace(void* dst, int dst_dtype, const void* src, int src_dtype, int64_t n, float alpha,
                                 void* stream) {
  if (n <= 0) return 0;
  MTVAF_REQUIRE(dst && src, "add_inplace: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(n, 256);
  if (dst_dtype == MTVAF_BF16 && src_dtype == MTVAF_BF16)
    add_inplace_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, n, alpha);
  else if (dst_dtype == MTVAF_BF16)
    add_inplace_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((__nv_bfloat16*)dst, (const float*)src, n, alpha);
  else if (src_dtype == MTVAF_BF16)
    add_inplace_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((float*)dst, (const __nv_bfloat16*)src, n, alpha);
  else
    add_inplace_kernel<float, float><<<g, 256, 0, st>>>((float*)dst, (const float*)src, n, alpha);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_colsum(const void* dy, int64_t ld, int dtype, int M, int N, float* db, void* stream) {
  MTVAF_REQUIRE(dy && db && M > 0 && N > 0, "colsum: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int esz = dtype == MTVAF_BF16 ? 2 : 4;
  const bool vec = N % 8 == 0 && (ld * esz) % 16 == 0 && reinterpret_cast<uintptr_t>(dy) % 16 == 0;
  const int cols_per_block = vec ? 256 : 64;
  const int col_blocks = (N + cols_per_block - 1) / cols_per_block;
  int row_blocks = (sm_count() * 8 + col_blocks - 1) / col_blocks;
  int rows_per_block = (M + row_blocks - 1) / row_blocks;
  const int gran = vec ? 32 : 8;
  rows_per_block = ((rows_per_block + gran - 1) / gran) * gran;
  if (rows_per_block < gran) rows_per_block = gran;
  row_blocks = (M + rows_per_block - 1) / rows_per_block;
  dim3 grid(col_blocks, row_blocks);
  if (vec) {
    if (dtype == MTVAF_BF16)
      colsum_vec_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, ld, M, N, rows_per_block, db);
    else
      colsum_vec_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, ld, M, N, rows_per_block, db);
  } else {
    if (dtype == MTVAF_BF16)
      colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, ld, M, N, rows_per_block, db);
    else
      colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, ld, M, N, rows_per_block, db);
  }
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_mean4_fwd(const void* x, void* y, int64_t rows, int W, int mode, int dtype, void* stream) {
  MTVAF_REQUIRE(x && y && rows > 0 && W > 0 && W % 4 == 0, "mean4_fwd: bad argument");
  const int g = grid_for(rows * W, 256);
  if (dtype == MTVAF_BF16)
    mean4_fwd_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, rows, W, mode);
  else
    mean4_fwd_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, rows, W, mode);
  MTVAF_LAUNCH_CHECK();
  return 0;
}
extern "C" int mtvaf_mean4_bwd_add(const float* dy, float* dx, int64_t rows, int W, int mode, void* stream) {
  MTVAF_REQUIRE(dy && dx && rows > 0 && W > 0 && W % 4 == 0, "mean4_bwd: bad argument");
  mean4_bwd_kernel<<<grid_for(rows * 4 * W, 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, rows, W, mode);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_rowscale(const void* x, const float* rowscale, void* y, int64_t M, int N, float alpha, int dtype,
                              void* stream) {
  if (M <= 0 || N <= 0) return 0;
  MTVAF_REQUIRE(x && rowscale && y, "rowscale: null pointer");
  const int g = grid_for(M * N, 256);
  if (dtype == MTVAF_BF16)
    rowscale_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rowscale,
                                                                        (__nv_bfloat16*)y, M, N, alpha);
  else
    rowscale_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)x, rowscale, (float*)y, M, N, alpha);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_scale_by_device_scalar(float* x, int64_t n, const float* scalar, void* stream) {
  if (n <= 0) return 0;
  MTVAF_REQUIRE(x && scalar, "scale_by_device_scalar: null pointer");
  scale_dev_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, scalar);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_mse_fwd_bwd(const float* norms, const float* labels, int64_t n, float* loss, float* dnorms,
                                 void* stream) {
  MTVAF_REQUIRE(norms && labels && loss && n > 0, "mse: bad argument");
  MTVAF_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream));
  int g = grid_for(n, 256);
  if (g > 64) g = 64;
  mse_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(norms, labels, n, 1.f / (float)n, loss, dnorms);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_combine_loss(const float* crf_nll_sum, int B, const float* prob_loss, float beta, int epoch,
                                  const float* img_losses, int n_img_losses, float alpha, float* out,
                                  int32_t* flag_out, void* stream) {
  MTVAF_REQUIRE(crf_nll_sum && out && B > 0, "combine_loss: bad argument");
  const float coef = beta * ldexpf(1.f, -epoch);
  combine_loss_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(crf_nll_sum, 1.f / (float)B, prob_loss, coef, img_losses,
                                                         img_losses ? n_img_losses : 0, alpha, out, flag_out);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                float grad_scale, void* bf16_copy, int zero_grad, const void* dyn, void* stream) {
  if (n <= 0) return 0;
  MTVAF_REQUIRE(param && grad && exp_avg && exp_avg_sq && (step >= 1 || dyn), "adamw: bad argument");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  auto al = [](const void* q, int a) { return reinterpret_cast<uintptr_t>(q) % a == 0; };
  const bool vec = al(param, 16) && al(grad, 16) && al(exp_avg, 16) && al(exp_avg_sq, 16) &&
                   (!bf16_copy || al(bf16_copy, 8));
  const AdamDyn* d = static_cast<const AdamDyn*>(dyn);
  if (vec)
    adamw_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale,
        (__nv_bfloat16*)bf16_copy, zero_grad, d);
  else
    adamw_scalar_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale,
        (__nv_bfloat16*)bf16_copy, zero_grad, d);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

// ---- row squared norms: out[r] = sum_c x[r][c]^2 (OneWordPSDProbe, probes/probe.py:74-78: the degenerate bmm of
// [T,1,r] x [T,r,1]).  One warp per row, 16-byte loads, shuffle reduction; cols % 8 == 0.
template <typename T>
__global__ void __launch_bounds__(256)
row_sqnorm_kernel(const T* __restrict__ x, long long ld, long long rows, int cols, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    const T* p = x + r * ld;
    float acc = 0.f;
    for (int c = lane * 8; c < cols; c += 256) {
      float v[8];
      Vec8<T>::load(p + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(v[j], v[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[r] = acc;
  }
}

extern "C" int mtvaf_row_sqnorm(const void* x, int64_t ld, int dtype, int64_t rows, int cols, float* out, void* stream) {
  MTVAF_REQUIRE(x && out && rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0, "row_sqnorm: bad argument");
  MTVAF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "row_sqnorm: x must be 16-byte aligned");
  const int grid = (int)std::min<long long>((rows + 7) / 8, (long long)sm_count() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MTVAF_BF16) row_sqnorm_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ld, rows, cols, out);
  else row_sqnorm_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ld, rows, cols, out);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

// ---- feature wire format: [B, E] + [B, n_aux, E] (fp32 / bf16) -> [1 + n_aux, B, E] (fp32 / bf16), 8 elements per thread
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
pack_features_kernel(const TI* __restrict__ img, long long img_ld, const TI* __restrict__ aux, long long aux_ld, int B,
                     int n_aux, long long E8, TO* __restrict__ out) {
  const long long per_img = (long long)B * E8;
  const long long total = per_img * (1 + n_aux);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i / per_img);
    const long long rem = i - (long long)j * per_img;
    const long long b = rem / E8, e = rem - b * E8;
    const TI* src = j == 0 ? img + b * img_ld + e * 8 : aux + b * aux_ld + ((long long)(j - 1) * E8 + e) * 8;
    float v[8];
    Vec8<TI>::load(src, v);
    Vec8<TO>::store(out + i * 8, v);
  }
}

extern "C" int mtvaf_pack_features(const void* images, int64_t img_ld, const void* aux_imgs, int64_t aux_ld, int in_dtype,
                                   int B, int n_aux, int64_t E, void* out, int out_dtype, void* stream) {
  MTVAF_REQUIRE(images && out && B > 0 && n_aux >= 0 && E > 0 && E % 8 == 0, "pack_features: bad argument");
  MTVAF_REQUIRE(img_ld % 8 == 0 && aux_ld % 8 == 0, "pack_features: sample strides must be multiples of 8 elements");
  MTVAF_REQUIRE(n_aux == 0 || aux_imgs, "pack_features: aux_imgs missing");
  MTVAF_REQUIRE(((reinterpret_cast<uintptr_t>(images) | reinterpret_cast<uintptr_t>(out) |
                  reinterpret_cast<uintptr_t>(aux_imgs)) & 15) == 0, "pack_features: pointers must be 16-byte aligned");
  const long long E8 = E / 8, total = (long long)B * E8 * (1 + n_aux);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
  cudaStream_t st = (cudaStream_t)stream;
#define MTVAF_PACK(TI, TO) \
  pack_features_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)images, img_ld, (const TI*)aux_imgs, aux_ld, B, n_aux, \
                                                     E8, (TO*)out)
  if (in_dtype == MTVAF_F32 && out_dtype == MTVAF_F32) MTVAF_PACK(float, float);
  else if (in_dtype == MTVAF_F32 && out_dtype == MTVAF_BF16) MTVAF_PACK(float, __nv_bfloat16);
  else if (in_dtype == MTVAF_BF16 && out_dtype == MTVAF_F32) MTVAF_PACK(__nv_bfloat16, float);
  else if (in_dtype == MTVAF_BF16 && out_dtype == MTVAF_BF16) MTVAF_PACK(__nv_bfloat16, __nv_bfloat16);
  else MTVAF_REQUIRE(false, "pack_features: unsupported dtype");
#undef MTVAF_PACK
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_adam_dyn_advance(void* dyn, float beta1, float beta2, int warmup_steps, int total_steps,
                                      void* stream) {
  MTVAF_REQUIRE(dyn && reinterpret_cast<uintptr_t>(dyn) % 8 == 0, "adam_dyn_advance: bad pointer");
  adam_dyn_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(static_cast<AdamDyn*>(dyn), beta1, beta2, warmup_steps,
                                                             total_steps);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_advance_step(uint64_t* dev_step, void* stream) {
  MTVAF_REQUIRE(dev_step, "advance_step: null pointer");
  advance_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(dev_step));
  MTVAF_LAUNCH_CHECK();
  return 0;
}
