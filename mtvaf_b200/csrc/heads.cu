// Heads and the small fused ops around the encoder:
//   * visual-prompt gates + prefix K/V packing   (models/bert_model.py:566-587)
//   * ANP softmax + KLDivLoss(batchmean)          (models/bert_model.py:553-554,560-561)
//   * psdProbe pseudo labels, bit-exact           (probes/constructLabel.py:11-29)
//   * TwoWordPSDProbe pairwise squared distances  (probes/probe.py:25-46)
//   * linear-chain CRF NLL (+grad) and Viterbi    (pytorch-crf semantics; bert_model.py:511,521)
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/mtvaf_b200.h"

namespace mtvaf {

// ================================================================ gates
// one block per (row = j*B + b, r): S2 = 2*hid columns x 4 splits, all layers
template <typename T>
__global__ void __launch_bounds__(256)
gate_fwd_kernel(const T* __restrict__ guids, const float* __restrict__ logits, int n_layers, int n_img, int B, int hid,
                T* __restrict__ kv_out, float* __restrict__ gates_out) {
  const int rowr = blockIdx.x;
  const int row = rowr >> 2, r = rowr & 3;
  const int j = row / B, b = row - j * B;
  const int S2 = 2 * hid, W = 4 * S2, P = 4 * n_img;
  extern __shared__ float sg[];                       // gates [n_layers][4]
  for (int l = threadIdx.x; l < n_layers; l += blockDim.x) {
    float v[4], mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x = logits[(long long)row * n_layers * 4 + l * 4 + i];
      x = x > 0.f ? x : 0.01f * x;                    // F.leaky_relu default slope
      v[i] = x;
      mx = fmaxf(mx, x);
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = expf(v[i] - mx); sum += v[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float g = v[i] / sum;
      sg[l * 4 + i] = g;
      if (r == 0) gates_out[(long long)row * n_layers * 4 + l * 4 + i] = g;
    }
  }
  __syncthreads();
  const T* src = guids + ((long long)row * 4 + r) * W;
  const long long kv_stride_l = 2LL * B * P * hid;    // per layer
  for (int c = threadIdx.x; c < S2; c += blockDim.x) {
    const float x0 = to_f<T>(src[c]), x1 = to_f<T>(src[S2 + c]), x2 = to_f<T>(src[2 * S2 + c]),
                x3 = to_f<T>(src[3 * S2 + c]);
    const int slot = c / hid, cc = c - slot * hid;
    T* dst = kv_out + ((long long)slot * B + b) * P * hid + (long long)(j * 4 + r) * hid + cc;
    for (int l = 0; l < n_layers; ++l) {
      // same accumulation order as the reference loop (:571-572): ((0 + g0 x0) + g1 x1) + ...
      float acc = 0.f + sg[l * 4 + 0] * x0;
      acc = acc + sg[l * 4 + 1] * x1;
      acc = acc + sg[l * 4 + 2] * x2;
      acc = acc + sg[l * 4 + 3] * x3;
      dst[l * kv_stride_l] = from_f<T>(acc);
    }
  }
}

// backward of the gate-weighted sums: one block per (row = j*B + b, r) as in forward.  Each thread keeps its
// columns' four split values and their gradient accumulators in registers across ALL layers, so the prompt
// (guids) is read once, d_guids is written once (plain store) and d_kv is streamed exactly once; the per-layer
// gate gradients are reduced warp-wise into shared memory and leave the block as one atomic per (layer, split).
//
// d_kv (600 MB at B=512, the kernel's whole cost) is STAGED through shared memory by 1-D bulk copies
// (cp.async.bulk): per layer the block's data are two contiguous runs of `hid` floats (the K and the V slot of its
// prefix row); a group of GL layers is requested at once on one mbarrier, two groups in flight.  With plain loads
// (6 scalar loads per thread and layer, no memory-level parallelism to speak of) the kernel sat at 1.2 TB/s.
template <typename T>
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const float* __restrict__ d_kv, const T* __restrict__ guids, const float* __restrict__ gates,
                int n_layers, int n_img, int B, int hid, int GL, float* __restrict__ d_guids,
                float* __restrict__ d_gates) {
  extern __shared__ __align__(128) uint8_t gsm[];
  constexpr int KMAX = 8;                             // columns per thread: 2*hid <= 2048
  const int rowr = blockIdx.x;
  const int row = rowr >> 2, r = rowr & 3;
  const int j = row / B, b = row - j * B;
  const int S2 = 2 * hid, W = 4 * S2, P = 4 * n_img;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stage = reinterpret_cast<float*>(gsm);                               // [2][GL][2 * hid]
  float* part = stage + 2 * GL * S2;                                          // [8 warps][n_layers * 4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + 8 * n_layers * 4);      // one per buffer
  const T* src = guids + ((long long)row * 4 + r) * W;
  float* dsrc = d_guids + ((long long)row * 4 + r) * W;
  const long long kv_stride_l = 2LL * B * P * hid;
  const int n_groups = (n_layers + GL - 1) / GL;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int g) {                            // one thread: layers [g*GL, ...) into buffer g & 1
    const int l0 = g * GL, nl = min(GL, n_layers - l0);
    float* dst = stage + (g & 1) * GL * S2;
    ptx::mbar_arrive_expect_tx(&bars[g & 1], (uint32_t)(nl * S2 * sizeof(float)));
    for (int l = 0; l < nl; ++l)
      for (int slot = 0; slot < 2; ++slot)
        ptx::bulk_load_1d(dst + l * S2 + slot * hid,
                          d_kv + (l0 + l) * kv_stride_l + ((long long)slot * B + b) * P * hid + (long long)(j * 4 + r) * hid,
                          (uint32_t)(hid * sizeof(float)), &bars[g & 1]);
  };
  if (threadIdx.x == 0) {
    issue(0);
    if (n_groups > 1) issue(1);
  }
  float x[KMAX][4], acc[KMAX][4];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = threadIdx.x + k * 256;
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[k][i] = c < S2 ? to_f<T>(src[i * S2 + c]) : 0.f; acc[k][i] = 0.f; }
  }
  const float* grow = gates + (long long)row * n_layers * 4;
  for (int g = 0; g < n_groups; ++g) {
    const int l0 = g * GL, nl = min(GL, n_layers - l0);
    const float* buf = stage + (g & 1) * GL * S2;
    ptx::mbar_wait(&bars[g & 1], (g >> 1) & 1);
    for (int ll = 0; ll < nl; ++ll) {
      const int l = l0 + ll;
      const float g0 = grow[l * 4 + 0], g1 = grow[l * 4 + 1], g2 = grow[l * 4 + 2], g3 = grow[l * 4 + 3];
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
      const float* dl = buf + ll * S2;                 // column c = slot * hid + cc: exactly the staged order
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = threadIdx.x + k * 256;
        if (c < S2) {
          const float d = dl[c];
          p0 += d * x[k][0]; p1 += d * x[k][1]; p2 += d * x[k][2]; p3 += d * x[k][3];
          acc[k][0] += g0 * d; acc[k][1] += g1 * d; acc[k][2] += g2 * d; acc[k][3] += g3 * d;
        }
      }
      p0 = warp_sum(p0); p1 = warp_sum(p1); p2 = warp_sum(p2); p3 = warp_sum(p3);
      if (lane == 0) {
        float* pp = part + warp * n_layers * 4 + l * 4;
        pp[0] = p0; pp[1] = p1; pp[2] = p2; pp[3] = p3;
      }
    }
    if (g + 2 < n_groups) {                            // this buffer is free again: request the group after next
      ptx::fence_proxy_async_smem();                   // generic-proxy reads before the async-proxy overwrite
      __syncthreads();
      if (threadIdx.x == 0) issue(g + 2);
    }
  }
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = threadIdx.x + k * 256;
    if (c < S2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dsrc[i * S2 + c] = acc[k][i];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n_layers * 4; idx += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w * n_layers * 4 + idx];
    atomicAdd(d_gates + (long long)row * n_layers * 4 + idx, s);
  }
}

// d(prompt) assembled in one pass, in the dtype of the GEMMs that consume it:
//   out[row, r, w] = d_guids[row, r, w] + ( d_gs[row, r*S + w % S] + dropout(d_gm[row, w]) ) / 4,   S = W / 4
// i.e. the gate path plus the backward of the two 4-way means of get_visual_prompt (bert_model.py:550,567);
// d_gm is the gradient of the ANP-head input BEFORE img_dropout (the mask is regenerated here).
template <typename TO, typename TM>
__global__ void __launch_bounds__(256)
prompt_grad_combine_kernel(const float* __restrict__ d_guids, const float* __restrict__ d_gs,
                           const TM* __restrict__ d_gm, long long rows, int W, TO* __restrict__ out,
                           uint32_t drop_thr, float drop_scale, unsigned long long seed_in,
                           const unsigned long long* __restrict__ step) {
  const unsigned long long seed = drop_thr ? step_seed(seed_in, step) : seed_in;
  const int S = W / 4;
  const long long n8 = rows * 4 * (W / 8);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const long long e = i * 8;
    const long long row = e / (4LL * W);
    const int rem = (int)(e - row * 4LL * W);
    const int r = rem / W, w = rem - r * W;
    float v[8];
    Vec8<float>::load(d_guids + e, v);
    if (d_gs) {
      float g[8];
      Vec8<float>::load(d_gs + row * W + r * S + (w % S), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += 0.25f * g[k];
    }
    if (d_gm) {
      float g[8];
      Vec8<TM>::load(d_gm + row * W + w, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float x = g[k];
        if (drop_thr) x = dropout_keep(seed, (unsigned long long)(row * W + w + k), drop_thr) ? x * drop_scale : 0.f;
        v[k] += 0.25f * x;
      }
    }
    Vec8<TO>::store(out + e, v);
  }
}

// softmax + leaky_relu backward on groups of 4
__global__ void gate_logit_bwd_kernel(const float* __restrict__ d_gates, const float* __restrict__ gates,
                                      const float* __restrict__ logits, long long n_groups,
                                      float* __restrict__ d_logits) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_groups) return;
  float g[4], dg[4], dot = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { g[k] = gates[i * 4 + k]; dg[k] = d_gates[i * 4 + k]; dot += g[k] * dg[k]; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float ds = g[k] * (dg[k] - dot);
    d_logits[i * 4 + k] = logits[i * 4 + k] > 0.f ? ds : 0.01f * ds;
  }
}

// ================================================================ softmax + KL(batchmean)
__global__ void __launch_bounds__(256)
softmax_kl_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ target, int B, int n,
                  float* __restrict__ loss_per_head, float* __restrict__ dlogits, float grad_scale) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int head = row / B, b = row - head * B;
  const float* x = logits + (long long)row * ld;
  const float* t = target + (long long)b * n;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, x[i]);
  mx = block_max(mx, red);
  float se = 0.f, st = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { se += expf(x[i] - mx); st += t[i]; }
  se = block_sum(se, red);
  st = block_sum(st, red);
  const float lse = mx + logf(se);
  float loss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float lq = x[i] - lse, ti = t[i];
    if (ti > 0.f) loss += ti * (logf(ti) - lq);       // xlogy convention: 0 * log 0 = 0
    if (dlogits) dlogits[(long long)row * ld + i] = (expf(lq) * st - ti) * grad_scale / (float)B;
  }
  loss = block_sum(loss, red);
  if (threadIdx.x == 0) atomicAdd(loss_per_head + head, loss / (float)B);
}

// ================================================================ probe pseudo labels (bit-exact)
// one block per sentence: bitonic sort of (value, index) = stable ascending sort, then thread 0 runs the
// sequential fp32 bucket scan of constructLabel.py:16-25, labels scattered back to original order.
__global__ void probe_labels_kernel(const float* __restrict__ norms, float* __restrict__ labels, int L, int n_pow2) {
  extern __shared__ unsigned char smraw[];
  float* val = reinterpret_cast<float*>(smraw);
  int* idx = reinterpret_cast<int*>(val + n_pow2);
  const float* v = norms + (long long)blockIdx.x * L;
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    val[i] = i < L ? v[i] : INFINITY;
    idx[i] = i < L ? i : 0x7fffffff;
  }
  __syncthreads();
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int p = i ^ jj;
        if (p > i) {
          const bool up = ((i & k) == 0);
          const float a = val[i], b = val[p];
          const int ia = idx[i], ib = idx[p];
          const bool a_gt_b = (a > b) || (a == b && ia > ib);
          if (a_gt_b == up) { val[i] = b; val[p] = a; idx[i] = ib; idx[p] = ia; }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    float lab = 0.f;
    for (int r = 0; r < L; ++r) {
      if (r == 0) lab = 1.f;
      else if (r == 1) lab = 2.f;
      else {
        const float x = val[r];
        const float d0 = fabsf(__fsub_rn(x, lab));
        const float d1 = fabsf(__fsub_rn(__fadd_rn(lab, 1.f), x));
        if (!(d0 < d1)) lab = __fadd_rn(lab, 1.f);
      }
      labels[(long long)blockIdx.x * L + idx[r]] = lab;
    }
  }
}

// ================================================================ TwoWord probe: explicit differences
template <typename T>
__global__ void __launch_bounds__(256)
pairwise_sqdist_kernel(const T* __restrict__ Tm, long long ld, int L, int R, float* __restrict__ dist) {
  __shared__ float A[16][33], Bt[16][33];
  const int b = blockIdx.z, i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc = 0.f;
  for (int r0 = 0; r0 < R; r0 += 32) {
    for (int e = threadIdx.x; e < 16 * 32; e += 256) {
      const int rr = e >> 5, cc = e & 31;
      const int gi = i0 + rr, gj = j0 + rr, gc = r0 + cc;
      A[rr][cc] = (gi < L && gc < R) ? to_f<T>(Tm[((long long)b * L + gi) * ld + gc]) : 0.f;
      Bt[rr][cc] = (gj < L && gc < R) ? to_f<T>(Tm[((long long)b * L + gj) * ld + gc]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 32; ++c) { const float d = A[ty][c] - Bt[tx][c]; acc = fmaf(d, d, acc); }
    __syncthreads();
  }
  if (i0 + ty < L && j0 + tx < L) dist[((long long)b * L + i0 + ty) * L + j0 + tx] = acc;
}

// ================================================================ CRF
// one warp per sequence; lane t < T owns tag t.  Left-aligned masks (length = sum(mask)), mask[:,0] == 1.
constexpr int CRF_MAX_T = 32;

__device__ __forceinline__ float warp_lse(float v, bool active) {
  const float x = active ? v : -INFINITY;
  const float mx = warp_max(x);
  const float e = active ? __expf(x - mx) : 0.f;
  return mx + __logf(warp_sum(e));
}

__global__ void __launch_bounds__(32)
crf_nll_kernel(const float* __restrict__ em, const long long* __restrict__ tags, const long long* __restrict__ mask,
               const float* __restrict__ start, const float* __restrict__ end, const float* __restrict__ trans, int L,
               int T, float* __restrict__ nll_sum, float* __restrict__ d_em, float* __restrict__ d_start,
               float* __restrict__ d_end, float* __restrict__ d_trans, float gs) {
  extern __shared__ float alpha[];                 // [L][T]
  __shared__ float tr[CRF_MAX_T * CRF_MAX_T];
  __shared__ float dtr[CRF_MAX_T * CRF_MAX_T];     // this sequence's transition gradient (lane t owns row t)
  const int b = blockIdx.x, t = threadIdx.x;
  const bool act = t < T;
  for (int i = t; i < T * T; i += 32) { tr[i] = trans[i]; dtr[i] = 0.f; }
  int len = 0;
  for (int i = t; i < L; i += 32) len += (mask[(long long)b * L + i] != 0);
  len = (int)(warp_sum((float)len) + 0.5f);
  __syncwarp();
  const float* e = em + (long long)b * L * T;
  const long long* tg = tags + (long long)b * L;
  // forward recursion
  float a = act ? start[t] + e[t] : -INFINITY;
  if (act) alpha[t] = a;
  for (int i = 1; i < len; ++i) {
    float mx = -INFINITY;
    float vals[CRF_MAX_T];
#pragma unroll 1
    for (int s = 0; s < T; ++s) {
      const float as = __shfl_sync(0xffffffffu, a, s);
      const float v = act ? as + tr[s * T + t] : -INFINITY;
      vals[s] = v;
      mx = fmaxf(mx, v);
    }
    float se = 0.f;
#pragma unroll 1
    for (int s = 0; s < T; ++s) se += act ? __expf(vals[s] - mx) : 0.f;
    a = act ? mx + __logf(se) + e[i * T + t] : -INFINITY;
    if (act) alpha[i * T + t] = a;
  }
  const float logz = warp_lse(a + (act ? end[t] : 0.f), act);
  // gold path score
  float score = 0.f;
  if (t == 0) {
    int prev = (int)tg[0];
    score = start[prev] + e[prev];
    for (int i = 1; i < len; ++i) {
      const int cur = (int)tg[i];
      score += tr[prev * T + cur] + e[i * T + cur];
      prev = cur;
    }
    score += end[prev];
    atomicAdd(nll_sum, logz - score);
  }
  if (!d_em) return;
  __syncwarp();
  // backward recursion: beta_i[t] = lse_u(trans[t,u] + em_{i+1}[u] + beta_{i+1}[u]); beta_{len-1} = end
  float beta = act ? end[t] : -INFINITY;
  for (int i = 0; i < L; ++i)
    if (act && i >= len) d_em[((long long)b * L + i) * T + t] = 0.f;
  for (int i = len - 1; i >= 0; --i) {
    const float al = act ? alpha[i * T + t] : -INFINITY;
    const float marg = act ? __expf(al + beta - logz) : 0.f;
    const int gold = (int)tg[i];
    if (act) d_em[((long long)b * L + i) * T + t] = (marg - (t == gold ? 1.f : 0.f)) * gs;
    if (i == 0 && act && d_start) atomicAdd(d_start + t, (marg - (t == gold ? 1.f : 0.f)) * gs);
    if (i == len - 1 && act && d_end) atomicAdd(d_end + t, (marg - (t == gold ? 1.f : 0.f)) * gs);
    if (i == 0) break;
    // pairwise marginals for transition (s -> u) at step i: alpha_{i-1}[s] + tr[s,u] + em_i[u] + beta_i[u] - logz
    const float w = act ? e[i * T + t] + beta : -INFINITY;     // indexed by u = lane
    const float ap = act ? alpha[(i - 1) * T + t] : -INFINITY; // indexed by s = lane
    const int gprev = (int)tg[i - 1];
    float nb_mx = -INFINITY;
    float vals[CRF_MAX_T];
#pragma unroll 1
    for (int u = 0; u < T; ++u) {
      const float wu = __shfl_sync(0xffffffffu, w, u);
      const float v = act ? tr[t * T + u] + wu : -INFINITY;    // lane = s
      vals[u] = v;
      nb_mx = fmaxf(nb_mx, v);
      if (act && d_trans) {
        const float pm = __expf(ap + v - logz);
        dtr[t * T + u] += (pm - ((t == gprev && u == gold) ? 1.f : 0.f)) * gs;   // one atomic per entry at the end
      }
    }
    float se = 0.f;
#pragma unroll 1
    for (int u = 0; u < T; ++u) se += act ? __expf(vals[u] - nb_mx) : 0.f;
    beta = act ? nb_mx + __logf(se) : -INFINITY;
  }
  if (d_trans) {
    __syncwarp();
    for (int i = t; i < T * T; i += 32) atomicAdd(d_trans + i, dtr[i]);
  }
}

__global__ void __launch_bounds__(32)
crf_decode_kernel(const float* __restrict__ em, const long long* __restrict__ mask, const float* __restrict__ start,
                  const float* __restrict__ end, const float* __restrict__ trans, int L, int T,
                  long long* __restrict__ best, long long* __restrict__ lengths) {
  extern __shared__ unsigned char hist[];          // [L][T] back-pointers
  __shared__ float tr[CRF_MAX_T * CRF_MAX_T];
  const int b = blockIdx.x, t = threadIdx.x;
  const bool act = t < T;
  for (int i = t; i < T * T; i += 32) tr[i] = trans[i];
  int len = 0;
  for (int i = t; i < L; i += 32) len += (mask[(long long)b * L + i] != 0);
  len = (int)(warp_sum((float)len) + 0.5f);
  __syncwarp();
  const float* e = em + (long long)b * L * T;
  float sc = act ? start[t] + e[t] : -INFINITY;
  for (int i = 1; i < len; ++i) {
    float bestv = -INFINITY;
    int bests = 0;
    for (int s = 0; s < T; ++s) {
      const float ps = __shfl_sync(0xffffffffu, sc, s);
      const float v = ps + (act ? tr[s * T + t] : 0.f);
      if (v > bestv) { bestv = v; bests = s; }     // strict >: first maximum wins (torch.max semantics)
    }
    if (act) hist[i * T + t] = (unsigned char)bests;
    sc = act ? bestv + e[i * T + t] : -INFINITY;
  }
  sc = act ? sc + end[t] : -INFINITY;
  // argmax over tags, lowest index on ties
  float bv = sc;
  int bi = t;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  __syncwarp();
  if (t == 0) {
    long long* out = best + (long long)b * L;
    int cur = bi;
    for (int i = len - 1; i >= 0; --i) {
      out[i] = cur;
      if (i > 0) cur = hist[i * T + cur];
    }
    for (int i = len; i < L; ++i) out[i] = -1;
    lengths[b] = len;
  }
}

}  // namespace mtvaf

using namespace mtvaf;

extern "C" int mtvaf_gate_fwd(const void* guids, const float* gate_logits, int n_layers, int n_img, int B, int hid,
                              void* kv_out, float* gates_out, int dtype, void* stream) {
  MTVAF_REQUIRE(guids && gate_logits && kv_out && gates_out, "gate_fwd: null argument");
  MTVAF_REQUIRE(n_layers > 0 && n_img > 0 && B > 0 && hid > 0, "gate_fwd: bad shape");
  const int blocks = n_img * B * 4;
  const size_t sm = (size_t)n_layers * 4 * sizeof(float);
  if (dtype == MTVAF_BF16)
    gate_fwd_kernel<__nv_bfloat16><<<blocks, 256, sm, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)guids, gate_logits, n_layers, n_img, B, hid, (__nv_bfloat16*)kv_out, gates_out);
  else
    gate_fwd_kernel<float><<<blocks, 256, sm, (cudaStream_t)stream>>>((const float*)guids, gate_logits, n_layers,
                                                                     n_img, B, hid, (float*)kv_out, gates_out);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_gate_bwd(const float* d_kv, const void* guids, const float* gate_logits, const float* gates,
                              int n_layers, int n_img, int B, int hid, float* d_guids, float* d_gates_scratch,
                              float* d_gate_logits, int dtype, void* stream) {
  MTVAF_REQUIRE(d_kv && guids && gate_logits && gates && d_guids && d_gates_scratch && d_gate_logits,
                "gate_bwd: null argument");
  MTVAF_REQUIRE(n_layers > 0 && n_img > 0 && B > 0 && hid > 0 && 2 * hid <= 2048, "gate_bwd: bad shape (hid <= 1024)");
  MTVAF_REQUIRE(hid % 4 == 0 && (reinterpret_cast<uintptr_t>(d_kv) & 15) == 0,
                "gate_bwd: d_kv must be 16-byte aligned with hid %% 4 == 0 (bulk copies)");
  const int blocks = n_img * B * 4;
  // layers per staged group: two groups (double buffer) of GL layers x 2 slots x hid floats, ~36 KB per group so that
  // three blocks share an SM
  int GL = (36 * 1024) / (2 * hid * (int)sizeof(float));
  if (GL < 1) GL = 1;
  if (GL > n_layers) GL = n_layers;
  const size_t sm = (size_t)2 * GL * 2 * hid * sizeof(float) + (size_t)8 * n_layers * 4 * sizeof(float) + 16;
  cudaStream_t st = (cudaStream_t)stream;
  static bool set = false;
  if (!set) {
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(gate_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    MTVAF_CHECK_CUDA(cudaFuncSetAttribute(gate_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    set = true;
  }
  MTVAF_REQUIRE(sm <= 100 * 1024, "gate_bwd: shared memory");
  if (dtype == MTVAF_BF16)
    gate_bwd_kernel<__nv_bfloat16><<<blocks, 256, sm, st>>>(d_kv, (const __nv_bfloat16*)guids, gates, n_layers, n_img,
                                                            B, hid, GL, d_guids, d_gates_scratch);
  else
    gate_bwd_kernel<float><<<blocks, 256, sm, st>>>(d_kv, (const float*)guids, gates, n_layers, n_img, B, hid, GL,
                                                    d_guids, d_gates_scratch);
  MTVAF_LAUNCH_CHECK();
  const long long groups = (long long)n_img * B * n_layers;
  gate_logit_bwd_kernel<<<(int)((groups + 255) / 256), 256, 0, st>>>(d_gates_scratch, gates, gate_logits, groups,
                                                                     d_gate_logits);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_prompt_grad_combine(const float* d_guids, const float* d_gs, const void* d_gm, int gm_dtype,
                                         float p_drop, uint64_t seed, int64_t rows, int W, void* out, int out_dtype,
                                         void* stream) {
  MTVAF_REQUIRE(d_guids && out && rows > 0 && W > 0 && W % 32 == 0, "prompt_grad_combine: bad argument (W %% 32 == 0)");
  MTVAF_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "prompt_grad_combine: bad dropout p");
  uint32_t thr = 0;
  float scale = 1.f;
  if (p_drop > 0.f && d_gm) {
    const double t = (double)p_drop * 4294967296.0;
    thr = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    scale = 1.f / (1.f - p_drop);
  }
  const long long n8 = rows * 4 * (W / 8);
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
#define MTVAF_PGC(TO_, TM_)                                                                                       \
  prompt_grad_combine_kernel<TO_, TM_><<<(int)blocks, 256, 0, st>>>(d_guids, d_gs, (const TM_*)d_gm, rows, W,    \
                                                                    (TO_*)out, thr, scale, seed, step_source())
  const bool ob = out_dtype == MTVAF_BF16, mb = gm_dtype == MTVAF_BF16;
  if (ob && mb) MTVAF_PGC(__nv_bfloat16, __nv_bfloat16);
  else if (ob) MTVAF_PGC(__nv_bfloat16, float);
  else if (mb) MTVAF_PGC(float, __nv_bfloat16);
  else MTVAF_PGC(float, float);
#undef MTVAF_PGC
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_softmax_kl_fwd_bwd(const float* logits, int64_t ld, const float* target, int rows, int B, int n,
                                        float* loss_per_head, float* dlogits, float grad_scale, void* stream) {
  MTVAF_REQUIRE(logits && target && loss_per_head && rows > 0 && B > 0 && n > 0 && rows % B == 0,
                "softmax_kl: bad argument");
  softmax_kl_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(logits, ld, target, B, n, loss_per_head, dlogits,
                                                            grad_scale);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_probe_labels(const float* norms, float* labels, int B, int L, void* stream) {
  MTVAF_REQUIRE(norms && labels && B > 0 && L > 0, "probe_labels: bad argument");
  int n2 = 1;
  while (n2 < L) n2 <<= 1;
  MTVAF_REQUIRE(n2 <= 4096, "probe_labels: L=%d too long", L);
  probe_labels_kernel<<<B, n2 < 256 ? (n2 < 32 ? 32 : n2) : 256, (size_t)n2 * 8, (cudaStream_t)stream>>>(norms, labels, L, n2);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

namespace mtvaf {
bool pairwise_tc_supported(const void* T, int64_t ld, int dtype, int R);            // pairwise_tc.cu
int pairwise_tc_launch(const float* T, int64_t ld, int B, int L, int R, float* dist, cudaStream_t st);
}  // namespace mtvaf
static int g_pairwise_impl = 0;
extern "C" int mtvaf_set_pairwise_impl(int impl) {
  MTVAF_REQUIRE(impl == 0 || impl == 1, "pairwise impl must be 0 (auto: tcgen05 Gram form when supported) or 1 (SIMT explicit differences)");
  g_pairwise_impl = impl;
  return 0;
}

extern "C" int mtvaf_pairwise_sqdist(const void* T, int64_t ld, int dtype, int B, int L, int R, float* dist,
                                     void* stream) {
  MTVAF_REQUIRE(T && dist && B > 0 && L > 0 && R > 0, "pairwise_sqdist: bad argument");
  // fp32 projected tokens with R % 64 == 0 (the probe ranks 384 / 512): Gram form on the tensor cores
  if (g_pairwise_impl == 0 && pairwise_tc_supported(T, ld, dtype, R))
    return pairwise_tc_launch((const float*)T, ld, B, L, R, dist, (cudaStream_t)stream);
  dim3 grid((L + 15) / 16, (L + 15) / 16, B);
  if (dtype == MTVAF_BF16)
    pairwise_sqdist_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)T, ld, L, R, dist);
  else
    pairwise_sqdist_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)T, ld, L, R, dist);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_crf_nll_fwd_bwd(const float* emissions, const int64_t* tags, const int64_t* mask,
                                     const float* start, const float* end, const float* trans, int B, int L, int T,
                                     float* nll_sum, float* d_emissions, float* d_start, float* d_end, float* d_trans,
                                     float grad_scale, void* stream) {
  MTVAF_REQUIRE(emissions && tags && mask && start && end && trans && nll_sum, "crf_nll: null argument");
  MTVAF_REQUIRE(T > 0 && T <= CRF_MAX_T && B > 0 && L > 0, "crf_nll: bad shape (T <= %d)", CRF_MAX_T);
  const size_t sm = (size_t)L * T * sizeof(float);
  MTVAF_REQUIRE(sm <= 40 * 1024, "crf_nll: L*T too large");
  crf_nll_kernel<<<B, 32, sm, (cudaStream_t)stream>>>(emissions, (const long long*)tags, (const long long*)mask, start,
                                                      end, trans, L, T, nll_sum, d_emissions, d_start, d_end, d_trans,
                                                      grad_scale);
  MTVAF_LAUNCH_CHECK();
  return 0;
}

extern "C" int mtvaf_crf_decode(const float* emissions, const int64_t* mask, const float* start, const float* end,
                                const float* trans, int B, int L, int T, int64_t* best_tags, int64_t* lengths,
                                void* stream) {
  MTVAF_REQUIRE(emissions && mask && start && end && trans && best_tags && lengths, "crf_decode: null argument");
  MTVAF_REQUIRE(T > 0 && T <= CRF_MAX_T && B > 0 && L > 0, "crf_decode: bad shape");
  const size_t sm = (size_t)L * T;
  MTVAF_REQUIRE(sm <= 40 * 1024, "crf_decode: L*T too large");
  crf_decode_kernel<<<B, 32, sm, (cudaStream_t)stream>>>(emissions, (const long long*)mask, start, end, trans, L, T,
                                                         (long long*)best_tags, (long long*)lengths);
  MTVAF_LAUNCH_CHECK();
  return 0;
}
