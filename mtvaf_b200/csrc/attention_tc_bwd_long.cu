// Backward of the prefix ("fusion") self-attention on tcgen05 for LONG text, 128 < L <= 512 (the roberta-large /
// long-auxiliary-text config and the reference's `--use_align` inputs of up to 500 tokens, MTVAF_training.py:250), any
// P <= 128.
//
// Same math as attention_tc_bwd.cu (flash-attention backward with the saved log-sum-exp,
// models/modeling_roberta.py:218-278); what differs is the blocking.  One CTA (512 threads) per (batch, head) item:
//   * the 128-query tiles are processed in GROUPS of two: Q and dO of a group stay resident in shared memory;
//   * per group the keys are walked in BLOCKS of <= 128: block 0 = the visual prefix (P8 rows), blocks 1.. = 128 text
//     keys each; K / V of one block at a time;
//   * per (key block, query tile): S = Q K^T and dP = dO V^T (128 x NB each) -> SIMT softmax -> bf16 P / dS in shared
//     memory -> dQ[tile] += dS K, dK[block] += dS^T Q, dV[block] += P^T dO;
//   * TMEM (512 columns): S [0,128) | dP [128,256) | dQ tile 0 [256,320) | dQ tile 1 [320,384) | dK [384,448) |
//     dV [448,512): dQ accumulates over the key blocks (final when its group ends), dK / dV over the query tiles of
//     the group.  With more than one group (L > 256) the later groups ADD their dK / dV to what the same thread of the
//     same CTA stored before (bf16 text rows, fp32 prefix rows): no atomics, no scratch buffer.
// Every step is separated by block barriers (no overlap between SIMT and MMA phases).
#include "attention_tc.cuh"

namespace mtvaf {
using namespace ptx;

namespace {

constexpr int kLongThreads = 512;
constexpr float kLog2eL = 1.4426950408889634f;
constexpr int LC_S = 0, LC_DP = 128, LC_DQ = 256, LC_DK = 384, LC_DV = 448;

struct LongSmem {
  size_t off_ds, off_p, off_q, off_do, off_k, off_v, off_mask, off_lse, off_d, off_exch, off_bar, total;
};
__host__ __device__ inline LongSmem long_layout() {
  LongSmem s;
  size_t o = 0;
  s.off_ds = o; o += 2 * 16384;          // dS of the current (block, tile): [128 q][128 keys] bf16, two 64-key chunks
  s.off_p = o;  o += 2 * 16384;
  s.off_q = o;  o += 2 * 16384;          // two query tiles
  s.off_do = o; o += 2 * 16384;
  s.off_k = o;  o += 16384;              // one key block
  s.off_v = o;  o += 16384;
  s.off_mask = o; o += 5 * 128 * sizeof(float);   // [block][key in block], additive mask * log2(e); <= 1 + 4 blocks
  s.off_lse = o;  o += 256 * sizeof(float);
  s.off_d = o;    o += 256 * sizeof(float);
  s.off_exch = o; o += 4 * 128 * sizeof(float);
  s.off_bar = o;  o += 64;
  s.total = o + 1024;
  return s;
}

__device__ __forceinline__ float long_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kLongThreads, 1)
attn_bwd_long_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmKp, const __grid_constant__ CUtensorMap tmVp,
                     const __grid_constant__ CUtensorMap tmdO, AttnTcArgs a, const float* __restrict__ lse,
                     const __nv_bfloat16* __restrict__ ctx, long long ld_ctx, __nv_bfloat16* __restrict__ dqkv,
                     long long ld_dqkv, float* __restrict__ dkp, float* __restrict__ dvp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const LongSmem lay = long_layout();
  uint8_t* sdS = smem + lay.off_ds;
  uint8_t* sP = smem + lay.off_p;
  uint8_t* sQ = smem + lay.off_q;
  uint8_t* sdO = smem + lay.off_do;
  uint8_t* sK = smem + lay.off_k;
  uint8_t* sV = smem + lay.off_v;
  float* sMask = reinterpret_cast<float*>(smem + lay.off_mask);
  float* sLse = reinterpret_cast<float*>(smem + lay.off_lse);
  float* sD = reinterpret_cast<float*>(smem + lay.off_d);
  float* sExch = reinterpret_cast<float*>(smem + lay.off_exch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.off_bar);
  uint64_t* bar_q = bars;          // Q and dO tiles of the item landed
  uint64_t* bar_kv = bars + 1;     // K and V of the current key block landed
  uint64_t* bar_s = bars + 2;      // S and dP in TMEM
  uint64_t* bar_g = bars + 3;      // gradient MMAs of the step retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;         // TMEM lane group / quarter of the 8-key units and of head dim
  const int row = quad * 32 + lane;                    // query row in its tile == key row in its block == TMEM lane
  const int H = a.nh * 64;
  const int n_items = a.B * a.nh;
  const int nq_all = (a.L + 127) / 128;                // query tiles (<= 4)
  const int nt = nq_all;                               // text key blocks
  const int n_groups = (nq_all + 1) / 2;               // query-tile groups of two
  const int n_blocks = nt + (a.P8 > 0 ? 1 : 0);
  const int first_text = a.P8 > 0 ? 1 : 0;

  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmKV); prefetch_tmap(&tmdO);
    if (a.P8 > 0) { prefetch_tmap(&tmKp); prefetch_tmap(&tmVp); }
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  // K / V rows a short prefix block does not load are still read by the MMAs, and P / dS columns a short block does not
  // write are read as never-stored accumulator rows: everything must be finite from the start (0 x NaN = NaN)
  for (int i = tid; i < (6 * 16384 + 2 * 16384) / 16; i += kLongThreads)
    *reinterpret_cast<uint4*>(smem + lay.off_ds + (size_t)i * 16) = make_uint4(0, 0, 0, 0);   // dS, P, Q, dO
  for (int i = tid; i < 2 * 16384 / 16; i += kLongThreads)
    *reinterpret_cast<uint4*>(sK + (size_t)i * 16) = make_uint4(0, 0, 0, 0);                  // K, V (contiguous)
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t aMask = smem_u32(sMask), aPs = smem_u32(sP), adSs = smem_u32(sdS);   // explicit shared-space accesses

  const bool has_prefix = a.P8 > 0;
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
  const float sc2 = a.scale * kLog2eL;
  const int row8 = row & 7;
  const uint32_t prow_off = (row >> 3) * 1024 + row8 * 128;
  const int dcol = part * 16;
  const float log2_ds = a.drop_thr ? log2f(a.drop_scale) : 0.f;
  const float ds_coef = a.scale / a.drop_scale;        // dS = P' * (scale / drop_scale) * (dP' - D)

  uint32_t n_kv = 0, n_step = 0, n_q = 0;              // completed phases of bar_kv / (bar_s, bar_g) / bar_q: same in every thread
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / a.nh, h = item - b * a.nh;
   for (int qg = 0; qg < n_groups; ++qg) {
    const int q0 = qg * 256;                             // first query row of the group
    const int nq = min(2, nq_all - qg * 2);              // query tiles in this group
    // ---- group prologue: Q / dO tiles by TMA; key masks, log-sum-exp and D_q = rowsum(dO o O) into shared memory
    __syncwarp();                                    // elect.sync needs the whole warp
    if (warp == 0 && elect_one()) {
      mbar_arrive_expect_tx(bar_q, (uint32_t)nq * 2u * 16384u);
      for (int qt = 0; qt < nq; ++qt) {
        tma_load_2d(sQ + qt * 16384, &tmQ, bar_q, h * 64, b * a.L + q0 + qt * 128);
        tma_load_2d(sdO + qt * 16384, &tmdO, bar_q, h * 64, b * a.L + q0 + qt * 128);
      }
    }
    for (int k = tid; qg == 0 && k < n_blocks * 128; k += kLongThreads) {
      const int kb = k >> 7, kk = k & 127;
      float m;
      if (has_prefix && kb == 0) m = (kk < a.P) ? 0.f : -INFINITY;
      else {
        const int t = (kb - first_text) * 128 + kk;
        m = (t < a.L) ? (a.key_mask[(long long)b * a.L + t] != 0 ? 0.f : -10000.0f * kLog2eL) : -INFINITY;
      }
      sMask[k] = m;
    }
    if (tid < 256) {
      const int q = q0 + tid;
      // +inf for rows past L: P = exp2(-inf) = 0
      sLse[tid] = (q < a.L) ? lse[((long long)b * a.nh + h) * a.L + q] * kLog2eL - log2_ds : INFINITY;
    }
    mbar_wait(bar_q, n_q & 1);
    ++n_q;
    for (int qt = 0; qt < nq; ++qt) {
      const int q = q0 + qt * 128 + row;
      float acc = 0.f;
      if (q < a.L) {
        const uint4* po = reinterpret_cast<const uint4*>(ctx + ((long long)b * a.L + q) * ld_ctx + h * 64 + dcol);
        const uint4 o0 = po[0], o1 = po[1];
        const uint8_t* pd = sdO + qt * 16384 + prow_off;
        const uint4 d0 = *reinterpret_cast<const uint4*>(pd + (((part * 2) ^ row8) << 4));
        const uint4 d1 = *reinterpret_cast<const uint4*>(pd + (((part * 2 + 1) ^ row8) << 4));
        const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const uint32_t ow[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 x = unpack_bf16x2(dw[j]), y = unpack_bf16x2(ow[j]);
          acc = fmaf(x.x, y.x, acc);
          acc = fmaf(x.y, y.y, acc);
        }
      }
      sExch[part * 128 + row] = acc;
      __syncthreads();
      if (part == 0) sD[qt * 128 + row] = (sExch[row] + sExch[128 + row]) + (sExch[256 + row] + sExch[384 + row]);
      __syncthreads();
    }

    // ---- key blocks
    for (int kb = 0; kb < n_blocks; ++kb) {
      const bool prefix_block = has_prefix && kb == 0;
      const int NB = prefix_block ? (a.P8 + 15) / 16 * 16 : 128;     // MMA N of S / dP, K extent of dQ
      const int units = NB >> 3;
      __syncwarp();                                    // elect.sync needs the whole warp
      if (warp == 0 && elect_one()) {
        // K / V of the previous block were last read by MMAs that have retired (bar_g awaited by everyone)
        if (prefix_block) {
          mbar_arrive_expect_tx(bar_kv, 2u * (uint32_t)a.P8 * 128u);
          for (int r = 0; r < a.P8; r += 8) {
            tma_load_2d(sK + r * 128, &tmKp, bar_kv, 0, (b * a.nh + h) * a.P + r);
            tma_load_2d(sV + r * 128, &tmVp, bar_kv, 0, (b * a.nh + h) * a.P + r);
          }
        } else {
          const int t0 = (kb - first_text) * 128;
          mbar_arrive_expect_tx(bar_kv, 2u * 16384u);
          for (int r = 0; r < 128; r += 64) {
            tma_load_2d(sK + r * 128, &tmKV, bar_kv, H + h * 64, b * a.L + t0 + r);
            tma_load_2d(sV + r * 128, &tmKV, bar_kv, 2 * H + h * 64, b * a.L + t0 + r);
          }
        }
      }
      for (int qt = 0; qt < nq; ++qt) {
        const int q = q0 + qt * 128 + row;
        __syncwarp();                                    // elect.sync needs the whole warp
        if (warp == 0 && elect_one()) {
          const uint32_t aQ = smem_u32(sQ + qt * 16384), aK = smem_u32(sK), adO = smem_u32(sdO + qt * 16384),
                         aV = smem_u32(sV);
          const uint32_t idesc = make_idesc_bf16(128, NB, false, false);
          if (qt == 0) mbar_wait(bar_kv, n_kv & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + LC_S, make_smem_desc_sw128(aQ + k * 32, 16, 1024),
                        make_smem_desc_sw128(aK + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + LC_DP, make_smem_desc_sw128(adO + k * 32, 16, 1024),
                        make_smem_desc_sw128(aV + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
          umma_commit(bar_s);
        }
        __syncwarp();
        const float lse2 = sLse[qt * 128 + row];
        const float dsum = sD[qt * 128 + row];
        const uint32_t rowkey =
            a.drop_thr ? attn_drop_rowkey(step_seed(a.seed, a.step), ((unsigned long long)b * a.nh + h) * a.L + q) : 0u;
        mbar_wait(bar_s, n_step & 1);
        __syncwarp();
        tc_fence_after();
        // ---- P and dS of this thread's (row, every 4th 8-key unit of the block)
        for (int u = part; u < units; u += 4) {
          const int c = u << 3;
          uint32_t rs[8], rd[8];
          tmem_ld_32x32b_x8(t_row + LC_S + c, rs);
          tmem_ld_32x32b_x8(t_row + LC_DP + c, rd);
          const float4 m0 = lds_f4(aMask + (kb * 128 + c) * 4);
          const float4 m1 = lds_f4(aMask + (kb * 128 + c) * 4 + 16);
          const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
          tmem_ld_wait();
          float p[8], dp[8], ds[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            p[j] = long_ex2(fmaf(__uint_as_float(rs[j]), sc2, mk[j] - lse2));
            dp[j] = __uint_as_float(rd[j]) * a.drop_scale;
            ds[j] = p[j] * ds_coef;
          }
          // reference key numbering for the dropout hash: prefix rows 0..P-1, then the text rows
          if (a.drop_thr)
            attn_drop_apply8(rowkey, prefix_block ? c : a.P + (kb - first_text) * 128 + c, a.drop_thr, p, dp);
#pragma unroll
          for (int j = 0; j < 8; ++j) ds[j] *= dp[j] - dsum;
          const uint32_t off = (u >> 3) * 16384 + prow_off + (((u & 7) ^ row8) << 4);
          uint4 w;
          w.x = pack_bf16x2(p[0], p[1]); w.y = pack_bf16x2(p[2], p[3]);
          w.z = pack_bf16x2(p[4], p[5]); w.w = pack_bf16x2(p[6], p[7]);
          sts_u4(aPs + off, w);
          w.x = pack_bf16x2(ds[0], ds[1]); w.y = pack_bf16x2(ds[2], ds[3]);
          w.z = pack_bf16x2(ds[4], ds[5]); w.w = pack_bf16x2(ds[6], ds[7]);
          sts_u4(adSs + off, w);
        }
        // key columns [NB, 128) of a short block keep older (finite) values: they only feed accumulator rows of
        // dK / dV that are never stored, and the dQ contraction stops at NB
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        __syncwarp();                                    // elect.sync needs the whole warp
        if (warp == 0 && elect_one()) {
          const uint32_t adS = smem_u32(sdS), aP = smem_u32(sP), aQ = smem_u32(sQ + qt * 16384),
                         adO = smem_u32(sdO + qt * 16384), aK = smem_u32(sK);
          // dQ[tile][q, d] += sum_key dS[q, key] K[key, d]
          const uint32_t idesc_q = make_idesc_bf16(128, 64, false, true);
          const int ksteps = NB / 16;
          for (int j = 0; j < ksteps; ++j)
            umma_f16_ss(tmem_base + LC_DQ + qt * 64, make_smem_desc_sw128(adS + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024),
                        make_smem_desc_sw128(aK + j * 2048, 8192, 1024), idesc_q, (kb > 0 || j > 0) ? 1u : 0u);
          // dK[block][key, d] += sum_q dS[q, key] Q[q, d] ; dV[block][key, d] += sum_q P[q, key] dO[q, d]
          const uint32_t idesc_t = make_idesc_bf16(128, 64, true, true);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_f16_ss(tmem_base + LC_DK, make_smem_desc_sw128(adS + j * 2048, 16384, 1024),
                        make_smem_desc_sw128(aQ + j * 2048, 8192, 1024), idesc_t, (qt > 0 || j > 0) ? 1u : 0u);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_f16_ss(tmem_base + LC_DV, make_smem_desc_sw128(aP + j * 2048, 16384, 1024),
                        make_smem_desc_sw128(adO + j * 2048, 8192, 1024), idesc_t, (qt > 0 || j > 0) ? 1u : 0u);
          umma_commit(bar_g);
        }
        __syncwarp();
        mbar_wait(bar_g, n_step & 1);
        __syncwarp();
        tc_fence_after();
        ++n_step;
      }
      ++n_kv;
      // ---- drain dK / dV of this key block: lane = key row of the block, this thread's 16 head-dim columns
      {
        const int ks = row;
        const int tx = (kb - first_text) * 128 + ks;
        const bool is_prefix = prefix_block && ks < a.P;
        const bool is_text = !prefix_block && tx < a.L;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          uint32_t r[16];
          __syncwarp();
          tmem_ld_32x32b_x16(t_row + (which ? LC_DV : LC_DK) + dcol, r);
          tmem_ld_wait();
          if (is_text) {
            uint4* o = reinterpret_cast<uint4*>(dqkv + ((long long)b * a.L + tx) * ld_dqkv + (which + 1) * H + h * 64 + dcol);
            if (qg > 0) {                        // add what this thread stored for the earlier query groups
#pragma unroll
              for (int v = 0; v < 2; ++v) {
                const uint4 old = o[v];
                const uint32_t ow[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = unpack_bf16x2(ow[j]);
                  r[v * 8 + 2 * j] = __float_as_uint(__uint_as_float(r[v * 8 + 2 * j]) + f.x);
                  r[v * 8 + 2 * j + 1] = __float_as_uint(__uint_as_float(r[v * 8 + 2 * j + 1]) + f.y);
                }
              }
            }
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
              w.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
              w.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
              w.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
              o[v] = w;
            }
          } else if (is_prefix) {
            float* o = (which ? dvp : dkp);
            if (o) {
              o += (((long long)b * a.nh + h) * a.P + ks) * 64 + dcol;
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                float4 acc = make_float4(__uint_as_float(r[v * 4 + 0]), __uint_as_float(r[v * 4 + 1]),
                                         __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
                if (qg > 0) {
                  const float4 old = *reinterpret_cast<const float4*>(o + v * 4);
                  acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
                }
                *reinterpret_cast<float4*>(o + v * 4) = acc;
              }
            }
          }
        }
      }
      // TMEM reads of dK / dV done before the next block's MMAs overwrite them; K / V smem free for the next block
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }

    // ---- drain dQ of the group's query tiles
    for (int qt = 0; qt < nq; ++qt) {
      const int q = q0 + qt * 128 + row;
      uint32_t r[16];
      __syncwarp();
      tmem_ld_32x32b_x16(t_row + LC_DQ + qt * 64 + dcol, r);
      tmem_ld_wait();
      if (q < a.L) {
        uint4* o = reinterpret_cast<uint4*>(dqkv + ((long long)b * a.L + q) * ld_dqkv + h * 64 + dcol);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
          w.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
          w.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
          w.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
          o[v] = w;
        }
      }
    }
    // all TMEM / shared-memory reads of this group done before the next group's / item's loads and MMAs
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
   }
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

bool attn_bwd_long_supported(const AttnTcArgs& a) {
  return a.L > 128 && a.L <= 512 && a.P8 <= 128 && long_layout().total <= 227 * 1024;
}

int attn_bwd_long_launch(const AttnTcArgs& a, const AttnTcMaps& m, const void* dctx, int64_t ld_dctx, const void* ctx,
                         int64_t ld_ctx, const float* lse, void* dqkv, int64_t ld_dqkv, float* dkp, float* dvp,
                         cudaStream_t st) {
  MTVAF_REQUIRE(ld_dqkv % 8 == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0,
                "attention_bwd(long): dqkv must be 16-byte aligned with ld %% 8 == 0");
  MTVAF_REQUIRE(ld_ctx % 8 == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0,
                "attention_bwd(long): ctx must be 16-byte aligned with ld %% 8 == 0");
  CUtensorMap tmdO;
  const uint64_t T = (uint64_t)a.B * a.L;
  const uint64_t H = (uint64_t)a.nh * 64;
  int rc = make_tmap_bf16_2d(&tmdO, dctx, H, T, ld_dctx, 64, 128);
  if (rc) return rc;
  const LongSmem lay = long_layout();
  static bool set = false;
  if (!set) {
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  const int n_items = a.B * a.nh;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  attn_bwd_long_kernel<<<grid, kLongThreads, lay.total, st>>>(m.q, m.kv, m.kp, m.vp, tmdO, a, lse,
                                                             (const __nv_bfloat16*)ctx, ld_ctx, (__nv_bfloat16*)dqkv,
                                                             ld_dqkv, dkp, dvp);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtvaf
