// TwoWordPSDProbe (probes/probe.py:25-46) on the tcgen05 tensor cores.
//
// The reference materialises the broadcast difference [B, L, L, R] and sums squares.  Here the pairwise squared
// distances come from the Gram matrix of the projected tokens, one 128 x 128 tile of (i, j) pairs per CTA:
//     d_ij = |t_i|^2 + |t_j|^2 - 2 <t_i, t_j>
// with fp32-class accuracy from bf16 tensor-core MMAs by the split  t = hi + lo  (hi = bf16(t), lo = bf16(t - hi)):
//     <t_i, t_j> = (hi_i.hi_j + lo_i.lo_j) + (hi_i.lo_j + lo_i.hi_j)      (t - hi - lo is below 2^-16 |t|)
// in three 128-column TMEM accumulators (fp32; bf16 x bf16 products are exact): hh+ll, hl, lh -- hl and lh separate so
// that the value of (i, j) and of (j, i) are sums of the same terms in the same order.
//   load : 256 threads read the fp32 rows (coalesced 16-byte loads), split them and write hi / lo in the K-major
//          SWIZZLE_128B operand layout; the row norms |t|^2 are accumulated in fp32 from the same registers;
//   MMA  : per 64-wide chunk of R: 4 x 4 tcgen05.mma (128 x 128 x 16), two-stage shared-memory ring so the next
//          chunk's loads run under the current chunk's MMAs;
//   out  : TMEM -> d tile in shared memory.  What the Gram form cannot give by itself is restored explicitly:
//          * the diagonal is EXACTLY 0 and the matrix EXACTLY symmetric (the reference's explicit differences are):
//            only tiles with i <= j are computed, the (j, i) block is the transposed copy of the same values, and
//            inside a diagonal tile entry (m, n) is read from (min, max);
//          * cancellation: where d_ij < 2^-10 (|t_i|^2 + |t_j|^2) -- near-duplicate tokens -- the pair is
//            recomputed from explicit fp32 differences (rare; the reference's accuracy for exactly those pairs).
// 2 R L^2 B flops on the tensor pipe; HBM traffic = T once per tile row/column + the fp32 [B, L, L] result.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/mtvaf_b200.h"

namespace mtvaf {
using namespace ptx;

namespace {

constexpr int kPwThreads = 256;
constexpr int kTile = 16384;                 // [128][64] bf16
constexpr int kStage = 4 * kTile;            // A_hi, A_lo, B_hi, B_lo
constexpr int kOutLd = 129;                  // fp32 d tile in smem, padded: conflict-free row AND column reads
constexpr int C_HH = 0, C_HL = 128, C_LH = 256;

// 32 fp32 values of one row (k = half*32 .. +32 of the chunk) -> four 16-byte pieces of hi and of lo; returns sum x^2
__device__ __forceinline__ float split_row_half(const float* __restrict__ src, bool valid, uint8_t* hi, uint8_t* lo,
                                                int row, int half) {
  const uint32_t rbase = (row >> 3) * 1024 + (row & 7) * 128;
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (valid) {
      a = *reinterpret_cast<const float4*>(src + q * 8);
      b = *reinterpret_cast<const float4*>(src + q * 8 + 4);
    }
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float r[8];
    uint32_t h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
      const float2 hf = unpack_bf16x2(h[j]);
      r[2 * j] = x[2 * j] - hf.x;            // exact in fp32
      r[2 * j + 1] = x[2 * j + 1] - hf.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(x[j], x[j], ss);
    const int piece = half * 4 + q;
    const uint32_t off = rbase + ((piece ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]),
                                                     pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
  }
  return ss;
}

__global__ void __launch_bounds__(kPwThreads, 1)
pairwise_gram_tc_kernel(const float* __restrict__ T, long long ld, int L, int R, int n_tiles,
                        float* __restrict__ dist) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage0 = smem;                                    // 2 x kStage operand ring; reused as the d tile
  float* sNA = reinterpret_cast<float*>(smem + 2 * kStage);  // |t_i|^2 of the tile's rows
  float* sNB = sNA + 128;                                    // |t_j|^2 of the tile's columns
  uint64_t* bars = reinterpret_cast<uint64_t*>(sNB + 128);   // one per stage
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // tile pair (ti <= tj) of sequence b
  const int pairs = n_tiles * (n_tiles + 1) / 2;
  const int b = blockIdx.x / pairs;
  int pr = blockIdx.x - b * pairs, ti = 0;
  while (pr >= n_tiles - ti) { pr -= n_tiles - ti; ++ti; }
  const int tj = ti + pr;
  const bool diag = ti == tj;
  const int i0 = ti * 128, j0 = tj * 128;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int row = tid >> 1, half = tid & 1;
  const float* rowA = T + ((long long)b * L + i0 + row) * ld + half * 32;
  const float* rowB = T + ((long long)b * L + j0 + row) * ld + half * 32;
  const bool okA = i0 + row < L, okB = j0 + row < L;
  float ssA = 0.f, ssB = 0.f;
  const int nk = R / 64;
  const uint32_t idesc = make_idesc_bf16(128, 128, false, false);

  for (int kc = 0; kc < nk; ++kc) {
    const int s = kc & 1;
    uint8_t* st = stage0 + s * kStage;
    if (kc >= 2) {                                           // the MMAs that read this stage (chunk kc-2) have retired
      mbar_wait(&bars[s], ((kc >> 1) - 1) & 1);
      tc_fence_after();
    }
    ssA += split_row_half(rowA + kc * 64, okA, st, st + kTile, row, half);
    if (!diag) ssB += split_row_half(rowB + kc * 64, okB, st + 2 * kTile, st + 3 * kTile, row, half);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    __syncwarp();                                    // elect.sync needs the whole warp
    if (warp == 0 && elect_one()) {
      const uint32_t aAh = smem_u32(st), aAl = aAh + kTile;
      const uint32_t aBh = diag ? aAh : aAh + 2 * kTile, aBl = diag ? aAl : aAh + 3 * kTile;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = (kc > 0 || k > 0) ? 1u : 0u;
        const uint64_t dAh = make_smem_desc_sw128(aAh + k * 32, 16, 1024), dAl = make_smem_desc_sw128(aAl + k * 32, 16, 1024);
        const uint64_t dBh = make_smem_desc_sw128(aBh + k * 32, 16, 1024), dBl = make_smem_desc_sw128(aBl + k * 32, 16, 1024);
        umma_f16_ss(tmem_base + C_HH, dAh, dBh, idesc, acc);
        umma_f16_ss(tmem_base + C_HH, dAl, dBl, idesc, 1u);        // lo.lo: symmetric too, shares the accumulator
        umma_f16_ss(tmem_base + C_HL, dAh, dBl, idesc, acc);
        umma_f16_ss(tmem_base + C_LH, dAl, dBh, idesc, acc);
      }
      umma_commit(&bars[s]);
    }
  }
  // row norms: the two halves of a row sit in adjacent lanes
  ssA += __shfl_xor_sync(0xffffffffu, ssA, 1);
  ssB += __shfl_xor_sync(0xffffffffu, ssB, 1);
  if (half == 0) {
    sNA[row] = ssA;
    sNB[row] = diag ? ssA : ssB;
  }
  // every MMA has retired (a commit covers all MMAs issued before it): TMEM is final, the operand ring is free
  mbar_wait(&bars[(nk - 1) & 1], ((nk - 1) >> 1) & 1);
  __syncthreads();
  tc_fence_after();

  // ---- TMEM -> d tile (fp32, [128][129]) in the freed operand ring
  float* sD = reinterpret_cast<float*>(stage0);
  {
    const int quad = warp & 3, ch = warp >> 2;               // TMEM lane quadrant / 64-column half
    const int m = quad * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float na = sNA[m];
    const int gi = i0 + m;
#pragma unroll 1
    for (int c0 = ch * 64; c0 < ch * 64 + 64; c0 += 16) {
      uint32_t hh[16], hl[16], lh[16];
      tmem_ld_32x32b_x16(t_row + C_HH + c0, hh);
      tmem_ld_32x32b_x16(t_row + C_HL + c0, hl);
      tmem_ld_32x32b_x16(t_row + C_LH + c0, lh);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = c0 + j, gj = j0 + n;
        const float g = __uint_as_float(hh[j]) + (__uint_as_float(hl[j]) + __uint_as_float(lh[j]));
        const float nb = sNB[n];
        float d = fmaxf((na + nb) - 2.f * g, 0.f);
        if (d < 9.765625e-4f * (na + nb) && gi < L && gj < L && gi != gj) {
          // near-duplicate rows: the Gram form has lost the digits -- explicit differences, as the reference computes
          const float* pa = T + ((long long)b * L + gi) * ld;
          const float* pb = T + ((long long)b * L + gj) * ld;
          float acc = 0.f;
          for (int r = 0; r < R; r += 4) {
            const float4 x = *reinterpret_cast<const float4*>(pa + r), y = *reinterpret_cast<const float4*>(pb + r);
            const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
            acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); acc = fmaf(d2, d2, acc); acc = fmaf(d3, d3, acc);
          }
          d = acc;
        }
        sD[m * kOutLd + n] = d;
      }
    }
  }
  tc_fence_before();
  __syncthreads();

  // ---- stores: block (ti, tj) row-major; block (tj, ti) as the transposed copy of the SAME values
  float* out = dist + (long long)b * L * L;
  for (int rr = warp; rr < 128; rr += kPwThreads / 32) {
    const int gi = i0 + rr;
    if (gi < L) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int n = c * 32 + lane, gj = j0 + n;
        if (gj < L) {
          float v;
          if (diag) v = (rr == n) ? 0.f : (rr < n ? sD[rr * kOutLd + n] : sD[n * kOutLd + rr]);
          else v = sD[rr * kOutLd + n];
          out[(long long)gi * L + gj] = v;
        }
      }
    }
    if (!diag) {
      const int gj = j0 + rr;                                 // row of the mirrored block
      if (gj < L) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int mm = c * 32 + lane, gi2 = i0 + mm;
          if (gi2 < L) out[(long long)gj * L + gi2] = sD[mm * kOutLd + rr];
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

bool pairwise_tc_supported(const void* T, int64_t ld, int dtype, int R) {
  return dtype == MTVAF_F32 && R >= 64 && R % 64 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(T) & 15) == 0;
}

int pairwise_tc_launch(const float* T, int64_t ld, int B, int L, int R, float* dist, cudaStream_t st) {
  const int n_tiles = (L + 127) / 128;
  const int pairs = n_tiles * (n_tiles + 1) / 2;
  const size_t smem = 1024 + 2 * (size_t)kStage + 256 * sizeof(float) + 64;
  static_assert(2 * kStage >= 128 * kOutLd * (int)sizeof(float), "d tile must fit in the operand ring");
  static bool set = false;
  if (!set) {
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(pairwise_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = true;
  }
  pairwise_gram_tc_kernel<<<B * pairs, kPwThreads, smem, st>>>(T, ld, L, R, n_tiles, dist);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtvaf
