// Backward of the prefix ("fusion") self-attention for SHORT text (L <= 64, P <= 16): TWO (batch, head) items per tile.
//
// Same pipeline as attention_tc_bwd_pipe.cu (control lane = TMA + MMA issue, 16 SIMT warps S / dP -> P / dS, 4 drain warps),
// but a 128-row tile of that kernel holds only L <= 64 query rows and a 128-key text tile only L <= 64 keys: at L = 64 the
// SIMT warps turn 128 x 144 scores into P / dS of which 64 x 80 exist.  Here rows [0, 64) are the queries of item 2i and rows
// [64, 128) those of item 2i + 1; key columns are [0, 64) text of a | [64, 128) text of b | [128, 128 + P8) prefix of a |
// [128 + P8, 128 + 2 P8) prefix of b.  S = [Q_a; Q_b] [K_a; K_b; Kp_a; Kp_b]^T and dP are each ONE MMA group whose
// off-diagonal blocks are never read; P and dS are block diagonal -- their off-diagonal pieces in shared memory are zeroed once
// per CTA and never written -- so dQ = dS K, dK = dS^T Q, dV = P^T dO and the transposed prefix gradients come out of the same
// MMAs as before with both items' rows side by side: the tensor pipe and every SIMT thread do work that is stored.
//   TMEM (512 columns): S [0,160) | dP [160,320) | dQ [320,384) | dK [384,448) | dV [448,512); the transposed prefix
//   gradients dK_p^T [0,32) | dV_p^T [32,64) ALIAS the S columns (free once P / dS are written): the drain warps read them
//   first and release them (bar_x) before the next item's S MMA is issued.
#include "attention_tc.cuh"
#include <cstdlib>

namespace mtvaf {
using namespace ptx;

namespace {

constexpr int kSimtThreads = 512;
constexpr int kDrainWarps = 4;
constexpr int kPipeThreads = kSimtThreads + 32 + kDrainWarps * 32;
constexpr float kLog2ePair = 1.4426950408889634f;
// TMEM columns
constexpr int TC_S = 0, TC_DP = 160, TC_DQ = 320, TC_DK = 384, TC_DV = 448, TC_DKP = 0, TC_DVP = 32;

struct PairSmem {
  int ns;                                             // key columns = K / V rows: 128 + 2 P8
  size_t k_stride;
  size_t off_ds, off_p, off_q, off_do, off_k, off_v, off_mask, off_exch, off_bar, total;
};

__host__ __device__ inline PairSmem pair_layout(int P8) {
  PairSmem s;
  s.ns = 128 + 2 * P8;
  s.k_stride = (size_t)s.ns * 128;                    // multiple of 1024 (P8 is a multiple of 8)
  size_t o = 0;
  // three 64-key chunks of [128][64] bf16: text a | text b | prefix a, prefix b (the rest of the chunk is never read)
  s.off_ds = o; o += 3 * 16384;
  s.off_p = o;  o += 3 * 16384;
  s.off_q = o;  o += 2 * 16384;
  s.off_do = o; o += 2 * 16384;
  s.off_k = o;  o += 2 * s.k_stride;
  s.off_v = o;  o += s.k_stride;
  s.off_mask = o; o += 2 * 160 * sizeof(float);
  s.off_exch = o; o += 2 * 128 * sizeof(float);
  s.off_bar = o; o += 128;
  s.total = o + 1024;
  return s;
}

__device__ __forceinline__ float pair_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void pair_simt_barrier() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__global__ void __launch_bounds__(kPipeThreads, 1)
attn_bwd_pair_kernel(const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmKp, const __grid_constant__ CUtensorMap tmVp,
                     const __grid_constant__ CUtensorMap tmdO, AttnTcArgs a, const float* __restrict__ lse,
                     const __nv_bfloat16* __restrict__ ctx, long long ld_ctx, __nv_bfloat16* __restrict__ dqkv,
                     long long ld_dqkv, float* __restrict__ dkp, float* __restrict__ dvp, float* __restrict__ dbias) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const PairSmem lay = pair_layout(a.P8);
  uint8_t* sdS = smem + lay.off_ds;
  uint8_t* sP = smem + lay.off_p;
  uint8_t* sQ0 = smem + lay.off_q;
  uint8_t* sdO0 = smem + lay.off_do;
  uint8_t* sK0 = smem + lay.off_k;
  uint8_t* sV = smem + lay.off_v;
  float* sMask0 = reinterpret_cast<float*>(smem + lay.off_mask);          // [2][256] additive mask * log2(e)
  float* sD0 = reinterpret_cast<float*>(smem + lay.off_exch);             // [2][128] D_q = rowsum(dO o O), by the drain warps
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.off_bar);
  uint64_t* bar_qk = bars;        // [2] Q + K tiles landed (TMA tx)
  uint64_t* bar_do = bars + 2;    // [2] dO tile landed
  uint64_t* bar_v = bars + 4;     //     V tile landed
  uint64_t* bar_s = bars + 5;     //     S and dP in TMEM            (tcgen05.commit)
  uint64_t* bar_p = bars + 6;     //     P and dS in smem, S/dP read (16 warp arrivals)
  uint64_t* bar_g = bars + 7;     //     gradients in TMEM           (tcgen05.commit)
  uint64_t* bar_o = bars + 8;     //     gradients drained           (4 drain-warp arrivals)
  uint64_t* bar_d = bars + 9;     // [2] D_q of the item in this buffer is in smem (4 drain-warp arrivals)
  uint64_t* bar_x = bars + 11;    //     prefix gradients (aliasing S) drained (2 drain-warp arrivals)
  uint64_t* bar_gx = bars + 12;   //     transposed prefix gradients in TMEM (tcgen05.commit, ahead of bar_g)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = a.nh * 64;
  const int n_bh = a.B * a.nh;
  const int n_items = (n_bh + 1) / 2;                  // pairs
  const int NT = 128;                                  // first prefix key column / smem row
  const int NS = lay.ns;                               // MMA N of S / dP: 128, 144 or 160
  const int cp = 2;                                    // chunk that holds the prefix columns
  // an odd item count leaves the last pair with one item: its second half re-reads the first (rows never stored)
  auto bh_of = [&](int item, int j) { const int bh = 2 * item + j; return bh < n_bh ? bh : 2 * item; };

  if (tid == 0) {
    prefetch_tmap(&tmKV); prefetch_tmap(&tmdO);
    if (a.P8 > 0) { prefetch_tmap(&tmKp); prefetch_tmap(&tmVp); }
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_p, 16);
    mbar_init(bar_g, 1);
    mbar_init(bar_o, kDrainWarps);
    mbar_init(&bar_d[0], kDrainWarps);
    mbar_init(&bar_d[1], kDrainWarps);
    mbar_init(bar_x, 2);
    mbar_init(bar_gx, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 16) tmem_alloc<512>(tmem_ptr);
  // P and dS start as zeros: their off-diagonal pieces stay that way for the life of the CTA (every K / V row is loaded)
  for (int i = tid; i < 2 * 3 * 1024; i += kPipeThreads) sts_u4(smem_u32(sdS) + i * 16, make_uint4(0, 0, 0, 0));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int first = blockIdx.x;

  if (warp == 16) {
    // ===================================== control: TMA + MMA issue ======================================
    if (elect_one()) {                                  // (not `lane == 0`: see attention_tc.cu)
      const uint32_t kv_bytes = (uint32_t)NS * 128u;
      auto load_qkdo = [&](int item, int buf) {
        uint8_t* q = sQ0 + buf * 16384;
        uint8_t* k = sK0 + buf * lay.k_stride;
        mbar_arrive_expect_tx(&bar_qk[buf], 16384u + kv_bytes);
        mbar_arrive_expect_tx(&bar_do[buf], 16384u);
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          const int bh = bh_of(item, j), b = bh / a.nh, h = bh - b * a.nh;
          tma_load_2d(q + j * 8192, &tmKV, &bar_qk[buf], h * 64, b * a.L);
          tma_load_2d(k + j * 8192, &tmKV, &bar_qk[buf], H + h * 64, b * a.L);
#pragma unroll 1
          for (int r = 0; r < a.P8; r += 8)
            tma_load_2d(k + (NT + j * a.P8 + r) * 128, &tmKp, &bar_qk[buf], 0, bh * a.P + r);
          tma_load_2d(sdO0 + buf * 16384 + j * 8192, &tmdO, &bar_do[buf], h * 64, b * a.L);
        }
      };
      auto load_v = [&](int item) {
        mbar_arrive_expect_tx(bar_v, kv_bytes);
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          const int bh = bh_of(item, j), b = bh / a.nh, h = bh - b * a.nh;
          tma_load_2d(sV + j * 8192, &tmKV, bar_v, 2 * H + h * 64, b * a.L);
#pragma unroll 1
          for (int r = 0; r < a.P8; r += 8)
            tma_load_2d(sV + (NT + j * a.P8 + r) * 128, &tmVp, bar_v, 0, bh * a.P + r);
        }
      };
      // ONE copy of everything this thread executes, as rolled loops: it runs alone, once per item, through code that the
      // other 20 warps' loops have long evicted from the 32 KB instruction cache -- with the loads and MMAs unrolled
      // (34 KB of code for this one thread) every 128-byte line of it was an L2 round trip (~85 cycles per MMA issued)
      const uint32_t idesc_s = make_idesc_bf16(128, NS, false, false);
      const uint32_t idesc_q = make_idesc_bf16(128, 64, false, true);
      const uint32_t idesc_t = make_idesc_bf16(128, 64, true, true);
      const uint32_t idesc_p = make_idesc_bf16(128, a.P8 > 8 ? 32 : 16, true, true);   // N >= 2 P8 (UMMA: N % 16 == 0)
      // The issuing thread is on the critical path of every item, so its instruction stream is kept short: the
      // descriptors of each operand are built once (per buffer) and only advanced -- the start-address field holds
      // addr >> 4, hence desc(base + off) = desc(base) + (off >> 4).
      // buffer 1 lives at a fixed byte distance from buffer 0: same trick for the per-buffer descriptors
      const uint64_t q_k0 = make_smem_desc_sw128(smem_u32(sQ0), 16, 1024);
      const uint64_t k_k0 = make_smem_desc_sw128(smem_u32(sK0), 16, 1024);
      const uint64_t o_k0 = make_smem_desc_sw128(smem_u32(sdO0), 16, 1024);
      const uint64_t k_mn0 = make_smem_desc_sw128(smem_u32(sK0), 8192, 1024);
      const uint64_t q_mn0 = make_smem_desc_sw128(smem_u32(sQ0), 8192, 1024);
      const uint64_t o_mn0 = make_smem_desc_sw128(smem_u32(sdO0), 8192, 1024);
      const uint64_t q_step = 16384 >> 4, k_step = lay.k_stride >> 4;
      const uint64_t dV_k = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
      const uint64_t dS_k = make_smem_desc_sw128(smem_u32(sdS), 16, 1024);
      const uint64_t dS_mn = make_smem_desc_sw128(smem_u32(sdS), 16384, 1024);
      const uint64_t P_mn = make_smem_desc_sw128(smem_u32(sP), 16384, 1024);
      const int ksteps = NS / 16;
      int il = -1;                                       // iteration -1 = prologue: only the first item's loads
#pragma unroll 1
      for (int item = first - (int)gridDim.x; item < n_items; item += gridDim.x, ++il) {
        const int next = item + gridDim.x;
        const bool live = il >= 0;
        const int buf = il & 1;
        const uint32_t ph = il & 1, ph2 = (il >> 1) & 1;
        const uint64_t q_k = q_k0 + buf * q_step, k_k = k_k0 + buf * k_step, o_k = o_k0 + buf * q_step;
        const uint64_t k_mn = k_mn0 + buf * k_step, q_mn = q_mn0 + buf * q_step, o_mn = o_mn0 + buf * q_step;
        if (live) {
          // ---- S = Q K^T, dP = dO V^T   (the S / dP columns are free: bar_p of the previous item was awaited
          //      before that item's gradient MMAs were issued -- except the S columns the previous item's transposed
          //      prefix gradients landed in: the drain warps release those first)
          mbar_wait(&bar_qk[buf], ph2);
          if (il > 0 && a.P8 > 0) mbar_wait(bar_x, ph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + TC_S, q_k + k * 2, k_k + k * 2, idesc_s, k > 0 ? 1u : 0u);
          mbar_wait(&bar_do[buf], ph2);
          mbar_wait(bar_v, ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + TC_DP, o_k + k * 2, dV_k + k * 2, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(bar_s);
          // ---- the other Q / K / dO buffers were last read by the previous item's gradient MMAs
          if (il > 0) mbar_wait(bar_g, ph ^ 1);
        }
        if (next < n_items) load_qkdo(next, (il + 1) & 1);
        if (live) mbar_wait(bar_s, ph);                  // dP retired: V is free
        if (next < n_items) load_v(next);
        if (!live) continue;
        // ---- gradients: need P / dS of this item and the previous item's gradients drained from TMEM
        mbar_wait(bar_p, ph);
        if (il > 0) mbar_wait(bar_o, ph ^ 1);
        tc_fence_after();
        if (a.P8 > 0) {
          // prefix keys FIRST (the next S MMA waits for their drain, which then runs under the text gradients), transposed:
          // dK_p^T[d, key] = sum_q Q[q, d] dS[q, key] ; dV_p^T[d, key] = sum_q dO[q, d] P[q, key]
          // (A = Q / dO as MN-major operands: rows 64..127 of the M = 128 tile read past the 64 real columns,
          //  finite or not: their output lanes are never stored).  These land in the S columns: P / dS are written
          //  (bar_p), so S is dead.
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_f16_ss(tmem_base + TC_DKP, q_mn + j * 128, dS_mn + cp * 1024 + j * 128, idesc_p, j > 0 ? 1u : 0u);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_f16_ss(tmem_base + TC_DVP, o_mn + j * 128, P_mn + cp * 1024 + j * 128, idesc_p, j > 0 ? 1u : 0u);
          umma_commit(bar_gx);
        }
        // dQ[q, d] = sum_key dS[q, key] K[key, d]   (K-major dS: 64-key chunks of 16 KB, 32 B per 16-key step)
#pragma unroll 5
        for (int j = 0; j < ksteps; ++j)
          umma_f16_ss(tmem_base + TC_DQ, dS_k + (j >> 2) * 1024 + (j & 3) * 2, k_mn + j * 128, idesc_q, j > 0 ? 1u : 0u);
        // text keys: dK[key, d] = sum_q dS[q, key] Q[q, d] ; dV[key, d] = sum_q P[q, key] dO[q, d]
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_f16_ss(tmem_base + TC_DK, dS_mn + j * 128, q_mn + j * 128, idesc_t, j > 0 ? 1u : 0u);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_f16_ss(tmem_base + TC_DV, P_mn + j * 128, o_mn + j * 128, idesc_t, j > 0 ? 1u : 0u);
        umma_commit(bar_g);
      }
    }
    __syncwarp();
  } else if (warp > 16) {
    // ============================================ drain warps =============================================
    // warp 17 + q' owns TMEM lanes [32 * quad, +32) with quad = warp & 3 (the lane quadrant a warp may access)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                  // query / key row == TMEM lane (head-dim index for the prefix)
    const int blk = quad >> 1, qrow = row & 63;        // which item of the pair / its query (= text key) row
    const bool row_in = qrow < a.L;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int row8 = row & 7;
    const uint32_t prow_off = (row >> 3) * 1024 + row8 * 128;
    // D_q = rowsum(dO o O) of `item` (whose dO tile lands in buffer `buf`): the O row comes straight from global
    // memory (128 contiguous bytes per lane), the dO row from the TMA tile.  Done here, one item AHEAD, because in
    // the softmax warps the prefetched O registers were spilled under the 80-register cap and the spill store
    // waited for the global load (profiles/r1_ncu_attn_v12.md: 5 % of samples + 18 % at the barrier behind it).
    auto compute_d = [&](int item, int buf, uint32_t parity) {
      const int bh = 2 * item + blk, b = bh / a.nh, h = bh - b * a.nh;
      mbar_wait(&bar_do[buf], parity);
      float acc = 0.f;
      if (row_in && bh < n_bh) {
        const uint4* po = reinterpret_cast<const uint4*>(ctx + ((long long)b * a.L + qrow) * ld_ctx + h * 64);
        const uint32_t pd = smem_u32(sdO0) + buf * 16384 + prow_off;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                 // two batches of four 16-byte loads (drain warps have slack)
          uint4 o[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) o[c] = __ldg(po + hf * 4 + c);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 dv = lds_u4(pd + (((hf * 4 + c) ^ row8) << 4));
            const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
            const uint32_t ow[4] = {o[c].x, o[c].y, o[c].z, o[c].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 x = unpack_bf16x2(dw[j]), y = unpack_bf16x2(ow[j]);
              acc = fmaf(x.x, y.x, acc);
              acc = fmaf(x.y, y.y, acc);
            }
          }
        }
      }
      sts_f32(smem_u32(sD0) + (buf * 128 + row) * 4, acc);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_d[buf]);
    };
    if (first < n_items) compute_d(first, 0, 0);
    int il = 0;
    for (int item = first; item < n_items; item += gridDim.x, ++il) {
      const int bh = 2 * item + blk, b = bh / a.nh, h = bh - b * a.nh;
      const bool row_ok = row_in && bh < n_bh;
      // the next item's D while the softmax warps work on this one (its buffer was last read two items ago)
      if (item + (int)gridDim.x < n_items) compute_d(item + gridDim.x, (il + 1) & 1, ((il + 1) >> 1) & 1);
      // prefix rows FIRST (they sit in the S columns the next item's S MMA is waiting for), transposed accumulators:
      // lane = head-dim index d (TMEM lanes 0..63), column = prefix key of item a [0, P8) | item b [P8, 2 P8)
      if (a.P8 > 0 && quad < 2) {                        // warp-uniform
        mbar_wait(bar_gx, il & 1);                       // committed ahead of the text gradients
        tc_fence_after();
        uint32_t rk[32], rv[32];
        tmem_ld_32x32b_x32(t_row + TC_DKP, rk);
        tmem_ld_32x32b_x32(t_row + TC_DVP, rv);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_x);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int bhj = 2 * item + j;
          if (bhj < n_bh) {
            const long long o = (long long)bhj * a.P * 64 + row;              // row == d here
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (k < a.P) {
                // column j * P8 + k: compile-time register index for P8 = 16, selected for P8 = 8
                const uint32_t vk = (a.P8 == 16 || j == 0) ? rk[j * 16 + k] : rk[(8 + k) & 31];
                const uint32_t vv = (a.P8 == 16 || j == 0) ? rv[j * 16 + k] : rv[(8 + k) & 31];
                if (dkp) dkp[o + k * 64] = __uint_as_float(vk);
                if (dvp) dvp[o + k * 64] = __uint_as_float(vv);
              }
          }
        }
      }
      mbar_wait(bar_g, il & 1);                        // this item's gradient MMAs have retired
      tc_fence_after();
      // dQ | dK | dV of the text rows: lane = row, 64 head-dim columns each = one full 128-byte line per lane
      // (rolled: one copy of the store + bias-gradient code instead of six keeps the kernel inside the instruction cache)
#pragma unroll 1
      for (int which = 0; which < 3; ++which) {
        const int tc = TC_DQ + which * 64;
        __nv_bfloat16* dst = dqkv + ((long long)b * a.L + qrow) * ld_dqkv + which * H + h * 64;
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + tc + hf * 32, r);
          tmem_ld_wait();
          if (row_ok) {
            uint4* o = reinterpret_cast<uint4*>(dst + hf * 32);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(r[v * 8 + 0]), __uint_as_float(r[v * 8 + 1]));
              w.y = pack_bf16x2(__uint_as_float(r[v * 8 + 2]), __uint_as_float(r[v * 8 + 3]));
              w.z = pack_bf16x2(__uint_as_float(r[v * 8 + 4]), __uint_as_float(r[v * 8 + 5]));
              w.w = pack_bf16x2(__uint_as_float(r[v * 8 + 6]), __uint_as_float(r[v * 8 + 7]));
              o[v] = w;
            }
          }
          if (dbias) {                                   // kernel-uniform
            // bias gradient of the fused QKV projection: column sums over this warp's 32 rows (transpose-reduce, 31
            // shuffles: after the stage with stride s a lane keeps the columns whose bit s equals its own), then one
            // fp32 reduction per column.  Replaces a separate pass over dqkv (the drain warps are off the critical path).
            float c[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) c[i] = row_ok ? __uint_as_float(r[i]) : 0.f;
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) {
              const bool up = (lane & st) != 0;
#pragma unroll
              for (int i = 0; i < st; ++i) {
                const float send = up ? c[i] : c[i + st];
                const float keep = up ? c[i + st] : c[i];
                c[i] = keep + __shfl_xor_sync(0xffffffffu, send, st);
              }
            }
            atomicAdd(dbias + which * H + h * 64 + hf * 32 + lane, c[0]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_o);
    }
  } else {
    // ============================================ SIMT warps ==============================================
    const int quad = warp & 3, part = warp >> 2;       // TMEM lane group / quarter of the columns
    const int row = quad * 32 + lane;                  // query row == TMEM lane
    const int blk = quad >> 1, qrow = row & 63;        // which item of the pair (warp-uniform) / its query row
    const bool row_in = qrow < a.L;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float sc2 = a.scale * kLog2ePair;
    const int row8 = row & 7;
    const uint32_t prow_off = (row >> 3) * 1024 + row8 * 128;
    const int units = 8 + (a.P8 >> 3);                 // per row: 8 text units + the item's prefix units
    const float log2_ds = a.drop_thr ? log2f(a.drop_scale) : 0.f;

    // key column c = tid: text row (c & 63) of item (c >> 6) for c < 128, then the prefix rows of item a, item b.
    // Only text columns inside the sequence depend on memory; what is PREFETCHED one item ahead is the raw key-mask word
    // and the raw log-sum-exp: nothing is computed from them until the next item uses them, so the loads have a whole
    // item to land (computing the mask value right after the load stalled every softmax warp on the long scoreboard once
    // per item -- 25 % of this kernel's stall samples, profiles/r2_ncu_attn_bwd_pair.md).
    const bool m_dyn = tid < NT && (tid & 63) < a.L;
    const float m_const = tid >= NS ? 0.f
                        : (tid < NT ? -INFINITY : (((tid - NT >= a.P8) ? tid - NT - a.P8 : tid - NT) < a.P ? 0.f : -INFINITY));
    auto fetch_mask_raw = [&](int item) -> long long {
      if (!m_dyn) return 1;
      const int b = bh_of(item, tid >> 6) / a.nh;
      return __ldg(a.key_mask + (long long)b * a.L + (tid & 63));
    };
    auto fetch_lse_raw = [&](int item) -> float {
      const int bh = 2 * item + blk;
      return (row_in && bh < n_bh) ? __ldg(lse + (long long)bh * a.L + qrow) : 0.f;
    };
    const unsigned long long seed_eff = a.drop_thr ? step_seed(a.seed, a.step) : 0ull;

    long long km_next = 1;
    float lse_next = 0.f;
    if (first < n_items) {
      km_next = fetch_mask_raw(first);
      lse_next = fetch_lse_raw(first);
    }
    int il = 0;
    int prev = -1;
    for (int item = first; item < n_items; item += gridDim.x, ++il) {
      const int bh = 2 * item + blk, b = bh / a.nh, h = bh - b * a.nh;
      const int buf = il & 1;
      const uint32_t ph = il & 1, ph2 = (il >> 1) & 1;
      const int next = item + gridDim.x;
      float* sMask = sMask0 + buf * 160;
      // ---- this item's prefetched scalars; the next item's travel while this one is processed
      if (tid < NS) sMask[tid] = m_dyn ? (km_next != 0 ? 0.f : -10000.0f * kLog2ePair) : m_const;
      const float lse2 = (row_in && bh < n_bh) ? lse_next * kLog2ePair - log2_ds : INFINITY;
      if (next < n_items) {
        km_next = fetch_mask_raw(next);
        lse_next = fetch_lse_raw(next);
      }
      // ---- P / dS in shared memory are free again once the previous item's gradient MMAs have retired (the
      //      drain warps take those gradients out of TMEM meanwhile)
      if (prev >= 0) mbar_wait(bar_g, ph ^ 1);
      pair_simt_barrier();                                    // publishes sMask
      // dS = P' (scale / drop_scale) (drop_scale dP_raw - D) = P' * scale * (dP_raw - D / drop_scale)
      mbar_wait(&bar_d[buf], ph2);                       // D_q of this item (drain warps, one item ahead)
      const float dsum_s = lds_f32(smem_u32(sD0) + (buf * 128 + row) * 4) / a.drop_scale;
      const float ds_c = a.scale;
      const uint32_t rowkey =
          a.drop_thr ? attn_drop_rowkey(seed_eff, ((unsigned long long)b * a.nh + h) * a.L + qrow) : 0u;

      mbar_wait(bar_s, ph);
      tc_fence_after();
      // ---- P and dS for this thread's (row, every 4th 8-key unit).  P' = P / (1-p) is born scaled (the dropout
      // scale rides in the exponent: lse2 carries -log2(scale)); dropped keys are zeroed in P' and in dP.
      const uint32_t aMask = smem_u32(sMask), aP = smem_u32(sP), adS = smem_u32(sdS);
      // local unit ul of this row's item -> S / dP column: text units 0..7, then the item's prefix units
      auto col_of = [&](int ul) { return ul < 8 ? (blk << 6) + (ul << 3) : NT + blk * a.P8 + ((ul - 8) << 3); };
      auto process_unit = [&](int ul, const uint32_t (&rs)[8], const uint32_t (&rd)[8]) {
        const int c = col_of(ul), u = c >> 3;
        const float4 m0 = lds_f4(aMask + c * 4);
        const float4 m1 = lds_f4(aMask + c * 4 + 16);
        const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        float p[8], dp[8], ds[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // masked / absent keys and rows past L give exp2(-inf) = 0 (S and dP are finite: padded K/V rows are 0)
          p[j] = pair_ex2(fmaf(__uint_as_float(rs[j]), sc2, mk[j] - lse2));
          dp[j] = __uint_as_float(rd[j]);                // raw dP' / drop_scale: the scale is folded into ds_c / dsum_s
          ds[j] = p[j] * ds_c;
        }
        // reference key numbering for the dropout hash: prefix rows 0..P-1, then the text rows
        if (a.drop_thr) attn_drop_apply8(rowkey, ul < 8 ? a.P + (ul << 3) : (ul - 8) << 3, a.drop_thr, p, dp);
#pragma unroll
        for (int j = 0; j < 8; ++j) ds[j] *= dp[j] - dsum_s;
        const uint32_t off = (u >> 3) * 16384 + prow_off + (((u & 7) ^ row8) << 4);
        uint4 w;
        w.x = pack_bf16x2(p[0], p[1]); w.y = pack_bf16x2(p[2], p[3]);
        w.z = pack_bf16x2(p[4], p[5]); w.w = pack_bf16x2(p[6], p[7]);
        sts_u4(aP + off, w);
        w.x = pack_bf16x2(ds[0], ds[1]); w.y = pack_bf16x2(ds[2], ds[3]);
        w.z = pack_bf16x2(ds[4], ds[5]); w.w = pack_bf16x2(ds[6], ds[7]);
        sts_u4(adS + off, w);
      };
      // (one rolled copy of the unit body: 2-3 units per thread; the other 15 warps cover the TMEM load latency)
#pragma unroll 1
      for (int u = part; u < units; u += 4) {
        uint32_t sa[8], da[8];
        tmem_ld_32x32b_x8(t_row + TC_S + col_of(u), sa);
        tmem_ld_32x32b_x8(t_row + TC_DP + col_of(u), da);
        tmem_ld_wait();
        process_unit(u, sa, da);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      prev = item;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

bool attn_bwd_pair_supported(const AttnTcArgs& a) {
  static const char* env = getenv("MTVAF_ATTN_BWD_PAIR");                  // experiment knob: 0 = off
  if (env && atoi(env) == 0) return false;
  return a.L <= 64 && a.P8 <= 16 && a.B * a.nh >= 2 && pair_layout(a.P8).total <= 227 * 1024;
}

int attn_bwd_pair_launch(const AttnTcArgs& a, const AttnTcMaps& m, const void* dctx, int64_t ld_dctx, const void* ctx,
                         int64_t ld_ctx, const float* lse, void* dqkv, int64_t ld_dqkv, float* dkp, float* dvp,
                         float* dbias, cudaStream_t st) {
  MTVAF_REQUIRE(ld_dqkv % 8 == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0,
                "attention_bwd(pair): dqkv must be 16-byte aligned with ld %% 8 == 0");
  MTVAF_REQUIRE(ld_ctx % 8 == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0,
                "attention_bwd(pair): ctx must be 16-byte aligned with ld %% 8 == 0");
  CUtensorMap tmdO;
  const uint64_t T = (uint64_t)a.B * a.L;
  const uint64_t H = (uint64_t)a.nh * 64;
  int rc = make_tmap_bf16_2d(&tmdO, dctx, H, T, ld_dctx, 64, 64);
  if (rc) return rc;
  const PairSmem lay = pair_layout(a.P8);
  static bool set = false;
  if (!set) {
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  const int n_items = (a.B * a.nh + 1) / 2;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  attn_bwd_pair_kernel<<<grid, kPipeThreads, lay.total, st>>>(m.kv, m.kp, m.vp, tmdO, a, lse,
                                                             (const __nv_bfloat16*)ctx, ld_ctx, (__nv_bfloat16*)dqkv,
                                                             ld_dqkv, dkp, dvp, dbias);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtvaf
