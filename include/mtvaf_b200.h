/* mtvaf_b200 -- C ABI of the B200-native (sm_100a) MTVAF hot path.
 *
 * The reference (MKMaS-GUET/MTVAF) has NO plugin/operator/FFI interface: the path sits behind plain
 * torch nn.Module classes.  The drop-in boundary is therefore the Python class
 * surface in mtvaf_b200/ (same class names, ctor and forward signatures as models/bert_model.py,
 * models/modeling_roberta.py, models/modeling_bert.py, probes/), and THIS header is the thin C ABI those
 * classes call through ctypes -- one entry point per ATen call-site group of the reference, cited
 * below as file:line of the reference code each one replaces.
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller (torch allocations) unless
 * noted; all functions are stream-ordered on `stream` (a cudaStream_t passed as void*), allocate
 * nothing, never synchronise and never throw: they return 0 on success or a negative code
 * (-1 bad argument, -2 CUDA error) with the message available from mtvaf_last_error().
 * dtype codes: MTVAF_F32 = 0 (parity mode: fp32 storage, fp32 SIMT math), MTVAF_BF16 = 1
 * (throughput mode: bf16 storage, tcgen05 tensor-core GEMMs with fp32 accumulation).
 */
#ifndef MTVAF_B200_H_
#define MTVAF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTVAF_ABI_VERSION 3   /* 2: MtvafEpilogue.colsum, mtvaf_attention_bwd_ex, mtvaf_set_sm_reserve
                               * 3: mtvaf_set_pairwise_impl, mtvaf_pack_features (tcgen05 TwoWord probe; feature wire format),
                               *    mtvaf_attention_fwd_ws / _workspace_bytes (long-text tcgen05 attention), mtvaf_row_sqnorm,
                               *    MTVAF_EPI_GELU_GRAD / MTVAF_EPI_MUL_AUX */
#define MTVAF_F32 0
#define MTVAF_BF16 1

/* ---- library ------------------------------------------------------------------------------- */
int mtvaf_abi_version(void);
const char* mtvaf_last_error(void);
/* number of CUDA kernels this library has launched so far in the process (statistics for benchmarks) */
uint64_t mtvaf_launch_count(void);
/* fills sm count and compute capability of the current device */
int mtvaf_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- GEMM family --------------------------------------------------------------------------- */
/* D[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) ).
 * Operand storage: `a_mn_major == 0`: A is [M,K] row-major (K contiguous, leading dim lda);
 *                  `a_mn_major == 1`: A is stored [K,M] (M contiguous, leading dim lda).
 *                  same for B with N in place of M.
 *   forward  Y = X W^T      (nn.Linear; modeling_roberta.py:202,219-220,296,365,379): A K-major, B K-major
 *   dgrad    dX = dY W      : A K-major, B MN-major
 *   wgrad    dW = dY^T X    : A MN-major, B MN-major
 */
enum {
  MTVAF_EPI_STORE = 0,      /* out = acc (+bias)                                                  */
  MTVAF_EPI_GELU = 1,       /* out2 = acc+bias (pre-activation, optional); out = gelu_erf(out2)   */
  MTVAF_EPI_TANH = 2,       /* out = tanh(acc+bias)                        (bert_model.py:446-454) */
  MTVAF_EPI_RESID = 3,      /* out = dropout(acc+bias) + aux               (modeling_roberta.py:296-298 minus LN) */
  MTVAF_EPI_ATOMIC_F32 = 4, /* out(f32) += acc   (split-K weight gradients into the grad bucket)   */
  MTVAF_EPI_MUL_DGELU = 5,  /* out = acc * gelu_erf'(aux)                  (backward of :365-366)  */
  MTVAF_EPI_MUL_DTANH = 6,  /* out = acc * (1 - aux^2)                                             */
  MTVAF_EPI_SQNORM = 7,     /* rowsum[m] += sum_n acc^2 ; out (optional) = acc   (probes/probe.py:74-78) */
  MTVAF_EPI_ROWSCALE = 8,   /* out = acc * rowscale[m]  (probe backward: 2 g_m T_m)                */
  MTVAF_EPI_GELU_GRAD = 9,  /* out = gelu(acc+bias); out2 (required) = gelu'(acc+bias): the derivative is computed in the
                             * FORWARD epilogue (which has issue slots to spare) and saved instead of the pre-activation */
  MTVAF_EPI_MUL_AUX = 10    /* out = acc * aux          (backward of :365-366 with aux = the saved gelu')            */
};

typedef struct MtvafEpilogue {
  int32_t mode;            /* MTVAF_EPI_*                                                         */
  int32_t out_dtype;       /* MTVAF_F32 / MTVAF_BF16 of `out` and `out2` (ATOMIC_F32: ignored)     */
  void* out;               /* [M, ldo]                                                            */
  int64_t ldo;
  const float* bias;       /* [N] fp32 or NULL                                                    */
  const void* aux;         /* residual / pre-activation / tanh output, same dtype as the operands */
  int64_t ld_aux;
  void* out2;              /* optional second output (pre-GELU), dtype out_dtype                  */
  int64_t ld_out2;
  float* rowvec;           /* SQNORM: [M] fp32 accumulated with atomics; ROWSCALE: [M] fp32 input  */
  float alpha;             /* scale applied to the accumulator first (1.0 = none)                 */
  float p_drop;            /* RESID: dropout probability (0 = off)                                */
  uint64_t seed;           /* RESID: dropout stream seed; element index = m * N + n               */
  float* colsum;           /* optional [N] fp32: += column sums of `out` as stored (the bias gradient of the layer
                              whose input gradient this GEMM produces, e.g. d(intermediate.dense.bias) from the
                              x GELU' epilogue).  Summed inside the TMA-staged tcgen05 epilogue from the staging box;
                              other kernel paths run mtvaf_colsum over `out` after the GEMM.  NULL = off.           */
} MtvafEpilogue;

/* bf16 operands, tcgen05.mma (kind::f16) with TMEM fp32 accumulators, TMA-fed, persistent.
 * Requirements: lda/ldb multiples of 8 elements, A/B base pointers 16-byte aligned.
 * `splits` > 1 partitions K over CTAs (only with MTVAF_EPI_ATOMIC_F32 / SQNORM is that meaningful). */
int mtvaf_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                    int M, int N, int K, const MtvafEpilogue* epi, int splits, void* stream);
/* 0 (default) = CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 tiles) whenever M >= 256, single-CTA
 * 128 x 256 tiles otherwise; 1 = single-CTA kernel only (A/B testing). */
int mtvaf_set_gemm_impl(int impl);
/* Leave `n_sms` SMs (rounded down to whole TPCs) out of the grids of the persistent kernels (GEMM, attention,
 * LayerNorm ...).  Data-parallel training sets this to the CTA budget of the gradient all-reduce while backward
 * runs: persistent kernels with a static tile schedule would otherwise wait for the SMs the collective holds and
 * run their share of tiles after everyone else.  0 (default) = use every SM.  Replaces nothing in the reference
 * (modules/parallel.py never overlaps communication). */
int mtvaf_set_sm_reserve(int n_sms);
/* fp32 operands, fp32 FFMA (parity mode; also used for skinny heads such as fc 768->11). */
int mtvaf_gemm_f32(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                   int M, int N, int K, const MtvafEpilogue* epi, int splits, void* stream);

/* out[m, n] = sum_k x[m, k] w[n, k] + bias[n] for skinny outputs N <= 48 (tag head fc 768->11, bert_model.py:510;
 * the 12 gate projectors as one [48, 6144] matrix, :566-569): warp-per-row, no split-K, bitwise reproducible. */
int mtvaf_skinny_linear_f32(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int M, int N,
                            int K, float* out, int64_t ldo, void* stream);

/* backward of the skinny layer (N <= 16).  dgrad: dx[m,k] = keep(seed, m*K+k)/(1-p) * sum_n dy[m,n] w[n,k], i.e. with
 * the dropout that precedes the tag head (bert_model.py:506) applied to the result, written as dx_dtype; when
 * `tanh_out` (same dtype / leading dim as dx) is given the result is also multiplied by 1 - tanh_out^2 (the
 * layer's input was tanh(dense(.)), bert_model.py:371-374);
 * wgrad: dw[n,k] += sum_m dy[m,n] x[m,k] (fp32 atomics), N*K/8 <= 1536. */
int mtvaf_skinny_linear_dgrad(const float* dy, int64_t lddy, const float* w, int64_t ldw, int M, int N, int K,
                              float p_drop, uint64_t seed, void* dx, int64_t lddx, int dx_dtype, const void* tanh_out,
                              void* stream);
int mtvaf_skinny_linear_wgrad(const float* dy, int64_t lddy, const float* x, int64_t ldx, int M, int N, int K,
                              float* dw, int64_t lddw, void* stream);

/* ---- elementwise / reductions --------------------------------------------------------------- */
int mtvaf_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);
int mtvaf_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream);
/* y[i] = keep(seed, i) ? x[i] / (1-p) : 0 -- the same counter-based mask the MTVAF_EPI_RESID epilogue uses
 * with element index i = m * N + n, so backward regenerates forward's mask (nn.Dropout :297,380, bert_model.py:506) */
int mtvaf_dropout_apply(const void* x, void* y, int64_t n, int dtype, float p_drop, uint64_t seed, void* stream);
/* dst[i] += alpha * src[i] with independent dtypes (gradient injection at hidden_states[7], residual joins) */
int mtvaf_add_inplace(void* dst, int dst_dtype, const void* src, int src_dtype, int64_t n, float alpha, void* stream);
/* y[m,n] = x[m,n] * alpha * rowscale[m]  (probe backward dT = 2 g_m T_m, probes/probe.py:74-78) */
int mtvaf_rowscale(const void* x, const float* rowscale, void* y, int64_t M, int N, float alpha, int dtype,
                   void* stream);
/* x[i] *= scalar[0], scalar on the device: scales head gradients by d(loss) without a host sync */
int mtvaf_scale_by_device_scalar(float* x, int64_t n, const float* scalar, void* stream);
/* db[n] += sum_m dY[m,n]   (bias gradients; fp32 atomics into the grad bucket) */
int mtvaf_colsum(const void* dy, int64_t ld, int dtype, int M, int N, float* db, void* stream);

/* ---- embeddings: RobertaEmbeddings.forward modeling_roberta.py:102-140 (+ :1706-1719),
 *      BertEmbeddings.forward modeling_bert.py:188-222 ------------------------------------------- */
/* kind 0 = roberta (position ids = cumsum(ids != pad) * (ids != pad) + pad, bit-exact int64),
 * kind 1 = bert (position ids = arange(L)).  Writes position ids (int64 [B,L]), the pre-LN sum is
 * not kept: mean/rstd [B*L] fp32 are saved for backward.  Dropout p on the output (0 = off). */
int mtvaf_embed_ln_fwd(const int64_t* input_ids, const int64_t* token_type_ids, const float* word_emb,
                       const float* pos_emb, const float* type_emb, const float* gamma, const float* beta,
                       float eps, int kind, int pad_idx, int B, int L, int H, int vocab, int max_pos, int n_types,
                       void* out, int out_dtype, int64_t* position_ids, float* mean, float* rstd, float p_drop,
                       uint64_t seed, void* stream);
/* backward: LN backward + scatter-add into the three tables (fp32 atomics; rows == padding_idx get
 * no gradient, as nn.Embedding(padding_idx) does: modeling_roberta.py:78,98-100). */
int mtvaf_embed_ln_bwd(const void* dout, int dtype, const int64_t* input_ids, const int64_t* token_type_ids,
                       const int64_t* position_ids, const float* word_emb, const float* pos_emb,
                       const float* type_emb, const float* gamma, const float* mean, const float* rstd, int kind,
                       int pad_idx, int B, int L, int H, float* d_word, float* d_pos, float* d_type, float* d_gamma,
                       float* d_beta, float p_drop, uint64_t seed, void* stream);

/* ---- LayerNorm (RobertaSelfOutput/RobertaOutput tails modeling_roberta.py:298,381) ------------ */
/* y = LN(z) * gamma + beta over rows of H; saves mean/rstd. z already holds dropout(dense)+residual
 * (fused into the GEMM epilogue MTVAF_EPI_RESID). */
int mtvaf_layernorm_fwd(const void* z, void* y, const float* gamma, const float* beta, float eps, int rows, int H,
                        int dtype, float* mean, float* rstd, void* stream);
/* dz = LN backward (dtype); d_gamma/d_beta accumulated with fp32 atomics.  Fused tail of the backward of
 * `LN(dropout(dense(x)) + residual)`: when p_drop > 0, `dd` (same dtype/shape as dz, required) receives
 * dropout_mask(seed, m*H+n) * dz / (1-p) -- the gradient entering the dense layer's GEMMs, same mask as the
 * forward MTVAF_EPI_RESID epilogue; when `d_bias` is non-NULL, d_bias[n] += sum_m dd[m,n] (dd == dz if p == 0),
 * i.e. the dense layer's bias gradient, so no separate dropout / column-sum pass is needed. */
int mtvaf_layernorm_bwd(const void* dy, const void* z, const float* gamma, const float* mean, const float* rstd,
                        int rows, int H, int dtype, void* dz, float* d_gamma, float* d_beta, void* dd,
                        float* d_bias, float p_drop, uint64_t seed, void* stream);

/* ---- prefix ("fusion") attention: RobertaSelfAttention.forward modeling_roberta.py:218-278 ---- */
/* qkv: [B*L, 3*nh*d] (Q | K | V column blocks, the fused QKV projection output), row stride ld_qkv.
 * kp, vp: prefix K/V [B, nh, P, d] (may be NULL when P == 0) -- keys/values are the prefix rows
 * FOLLOWED by the text rows (torch.cat at :221-222), one softmax over P+L keys.
 * key_mask: [B, L] int64 text attention mask (prefix columns are always visible, bert_model.py:490-492);
 * masked keys get the additive -10000.0 of modeling_roberta.py:1000.
 * ctx: [B*L, nh*d] (heads merged, :276-278).  lse: [B, nh, L] fp32 log-sum-exp saved for backward.
 * probs (optional, may be NULL): [B, nh, L, P+L] fp32 attention probabilities (output_attentions). */
int mtvaf_attention_fwd(const void* qkv, int64_t ld_qkv, const void* kp, const void* vp, int P,
                        const int64_t* key_mask, int B, int L, int nh, int d, void* ctx, int64_t ld_ctx, float* lse,
                        float* probs, int dtype, float p_drop, uint64_t seed, void* stream);
/* Same, with a caller-provided workspace (device memory, 256-byte aligned, mtvaf_attention_fwd_workspace_bytes(...)
 * bytes): long bf16 text whose keys do not fit one resident tile set (P + L > ~400, L <= 512) then runs on tcgen05 as two
 * key windows + a merge of the partial softmaxes instead of the SIMT kernel.  workspace may be NULL (= mtvaf_attention_fwd). */
int64_t mtvaf_attention_fwd_workspace_bytes(int B, int L, int nh, int d, int P, int dtype);
int mtvaf_attention_fwd_ws(const void* qkv, int64_t ld_qkv, const void* kp, const void* vp, int P,
                        const int64_t* key_mask, int B, int L, int nh, int d, void* ctx, int64_t ld_ctx, float* lse,
                        float* probs, int dtype, float p_drop, uint64_t seed, void* workspace, int64_t workspace_bytes, void* stream);
/* 0 (default) = tcgen05 kernels whenever dtype is bf16 and the shape fits (L <= 512, P <= 128; forward with P + L beyond
 *     one resident tile set needs the workspace of mtvaf_attention_fwd_ws), SIMT otherwise;
 * 1 = SIMT kernels only, 2 = tcgen05 kernels but the generic (non-pipelined) backward for L <= 128 (A/B testing).
 * Process-global. */
int mtvaf_set_attention_impl(int impl);
/* dqkv: [B*L, 3*nh*d]; dkp/dvp: [B, nh, P, d] fp32 gradient of the prefix (may be NULL);
 * dsum_scratch: [B, nh, L] fp32 workspace (rowsum(dO * O)). */
int mtvaf_attention_bwd(const void* dctx, int64_t ld_dctx, const void* qkv, int64_t ld_qkv, const void* kp,
                        const void* vp, int P, const int64_t* key_mask, const void* ctx, int64_t ld_ctx,
                        const float* lse, int B, int L, int nh, int d, void* dqkv, int64_t ld_dqkv, float* dkp,
                        float* dvp, float* dsum_scratch, int dtype, float p_drop, uint64_t seed, void* stream);
/* Same, plus d_bias_qkv (optional, may be NULL): fp32 [3*nh*d], += column sums of dqkv = the bias gradient of the fused
 * Q/K/V projection (autograd of modeling_roberta.py:202,219-220).  Produced inside the pipelined tcgen05 kernel
 * while its drain warps empty TMEM; other kernel paths run mtvaf_colsum over dqkv afterwards. */
int mtvaf_attention_bwd_ex(const void* dctx, int64_t ld_dctx, const void* qkv, int64_t ld_qkv, const void* kp,
                        const void* vp, int P, const int64_t* key_mask, const void* ctx, int64_t ld_ctx,
                        const float* lse, int B, int L, int nh, int d, void* dqkv, int64_t ld_dqkv, float* dkp,
                        float* dvp, float* dsum_scratch, int dtype, float p_drop, uint64_t seed, float* d_bias_qkv, void* stream);

/* ---- visual prompt gates: get_visual_prompt bert_model.py:566-587 ----------------------------- */
/* guids: [n_img, B, 4, 8*hid] MLP outputs (encoder_conv), each row viewed as 4 splits of 2*hid.
 * gate_logits: [n_img*B, n_layers*4] fp32 = all projectors[l] applied to the mode-1 mean (one GEMM).
 * gates_out:   [n_img*B, n_layers*4] fp32 = softmax(leaky_relu(logits)) per group of 4 (saved for bwd).
 * kv_out: [n_layers, 2, B, P*hid] with P = 4*n_img: for layer l the flat [P, hid] matrices whose
 * plain reshape(bsz, nh, -1, d) (NOT a head transpose, bert_model.py:585) is the prefix K (slot 0) / V
 * (slot 1): kv_out[l,s,b,(j*4+r)*hid + c] = sum_i gate[l,(j,b),i] * guids[j,b,r,i*2*hid + s*hid + c]. */
int mtvaf_gate_fwd(const void* guids, const float* gate_logits, int n_layers, int n_img, int B, int hid,
                   void* kv_out, float* gates_out, int dtype, void* stream);
/* d_kv: [n_layers, 2, B, P*hid] fp32.  d_guids (fp32 [n_img,B,4,8*hid]) is WRITTEN (the gate path's share of the
 * prompt gradient); d_gates_scratch [n_img*B, n_layers*4] fp32 must be zeroed by the caller; d_gate_logits same
 * shape (written). */
int mtvaf_gate_bwd(const float* d_kv, const void* guids, const float* gate_logits, const float* gates, int n_layers,
                   int n_img, int B, int hid, float* d_guids, float* d_gates_scratch, float* d_gate_logits,
                   int dtype, void* stream);
/* out[row, r, w] = d_guids[row, r, w] + ( d_gs[row, r*S + w % S] + dropout(d_gm[row, w]) ) / 4 with S = W/4:
 * the gate path plus the backward of both 4-way means of get_visual_prompt (bert_model.py:550,567), written once
 * in the dtype of the GEMMs that consume it.  d_gs (fp32 [rows, W], gradient of the mode-1 mean) and d_gm
 * ([rows, W] of gm_dtype, gradient of the ANP-head input after img_dropout, bert_model.py:551; the dropout mask
 * (p_drop, seed, element index row*W + w) is re-applied here) may each be NULL. */
int mtvaf_prompt_grad_combine(const float* d_guids, const float* d_gs, const void* d_gm, int gm_dtype, float p_drop,
                              uint64_t seed, int64_t rows, int W, void* out, int out_dtype, void* stream);
/* 4-way means of the prompt x [rows, 4, W] -> y [rows, W]:
 *   mode 0: y[row, w]       = mean_r x[row, r, w]                      (guids.mean(dim=1), bert_model.py:550)
 *   mode 1: y[row, r*S + c] = mean_i x[row, r, i*S + c], S = W/4       (stack(split).sum(0)/4, bert_model.py:567)
 * backward: dx (fp32 [rows,4,W]) += broadcast(dy) / 4 */
int mtvaf_mean4_fwd(const void* x, void* y, int64_t rows, int W, int mode, int dtype, void* stream);
int mtvaf_mean4_bwd_add(const float* dy, float* dx, int64_t rows, int W, int mode, void* stream);

/* ---- ANP heads: softmax + KLDivLoss(batchmean) bert_model.py:553-554,560-561 ------------------ */
/* logits [rows, ld] fp32, target [B, n] fp32 (row r uses target row r % B); loss_out[head] += ... with
 * head = r / B; dlogits (optional) = d loss / d logits * scale. */
int mtvaf_softmax_kl_fwd_bwd(const float* logits, int64_t ld, const float* target, int rows, int B, int n,
                             float* loss_per_head, float* dlogits, float grad_scale, void* stream);

/* ---- visual features: wire format -> GEMM operand (models/bert_model.py:536-539) ------------------ */
/* images [B, E] (samples `img_ld` elements apart), aux_imgs [B, n_aux, E] (samples `aux_ld` apart; both `in_dtype`:
 * fp32, or bf16 = the cached wire format of the frozen ResNet pyramid, E = 3840*2*2) -> out [1 + n_aux, B, E] in
 * `out_dtype`: the cat + view + aux permute + cast of get_visual_prompt's first lines in one pass.  E and the strides
 * % 8 == 0, 16-byte aligned pointers. */
int mtvaf_pack_features(const void* images, int64_t img_ld, const void* aux_imgs, int64_t aux_ld, int in_dtype, int B,
                        int n_aux, int64_t E, void* out, int out_dtype, void* stream);

/* ---- psdProbe: probes/probe.py:74-78, probes/constructLabel.py:11-29, probes/probe_trainModel.py:23-24 */
/* OneWordPSDProbe probes/probe.py:74-78: out[r] = sum_c x[r][c]^2 over the projected tokens x [rows, cols] (cols % 8 == 0):
 * the bf16 path stores T = x proj with the TMA-store epilogue and takes the norms from it in one HBM-bound pass. */
int mtvaf_row_sqnorm(const void* x, int64_t ld, int dtype, int64_t rows, int cols, float* out, void* stream);
/* bit-exact pseudo labels (stable sort + sequential fp32 scan) for norms [B, L] fp32 -> labels [B, L] fp32 */
int mtvaf_probe_labels(const float* norms, float* labels, int B, int L, void* stream);
/* loss[0] = mean((norms-labels)^2) ; dnorms (optional) = 2 (norms-labels) / (B L) */
int mtvaf_mse_fwd_bwd(const float* norms, const float* labels, int64_t n, float* loss, float* dnorms, void* stream);
/* TwoWordPSDProbe probes/probe.py:25-46: dist[b,i,j] = sum_r (T[b,i,r]-T[b,j,r])^2.  fp32 T with R % 64 == 0: Gram form
 * on the tcgen05 tensor cores (bf16 hi/lo split, fp32-class accuracy; diagonal exactly 0, exactly symmetric,
 * near-duplicate pairs recomputed from explicit differences); otherwise the SIMT explicit-difference kernel. */
int mtvaf_pairwise_sqdist(const void* T, int64_t ld, int dtype, int B, int L, int R, float* dist, void* stream);
/* 0 = auto, 1 = always the SIMT explicit-difference kernel (A/B testing; process-global) */
int mtvaf_set_pairwise_impl(int impl);

/* ---- linear-chain CRF (pytorch-crf semantics; call sites bert_model.py:464,511,521) ----------- */
/* emissions [B, L, T] fp32, tags [B, L] int64, mask [B, L] int64 (mask[:,0] must be 1).
 * nll_sum[0] += sum_b -(score - logZ); d_emissions (optional) = d(nll_mean)/d emissions;
 * d_start/d_end/d_trans (optional) accumulated with atomics, all scaled by grad_scale (1/B for mean). */
int mtvaf_crf_nll_fwd_bwd(const float* emissions, const int64_t* tags, const int64_t* mask, const float* start,
                          const float* end, const float* trans, int B, int L, int T, float* nll_sum,
                          float* d_emissions, float* d_start, float* d_end, float* d_trans, float grad_scale,
                          void* stream);
/* Viterbi: best_tags [B, L] int64 (positions >= length are -1), lengths [B] int64 */
int mtvaf_crf_decode(const float* emissions, const int64_t* mask, const float* start, const float* end,
                     const float* trans, int B, int L, int T, int64_t* best_tags, int64_t* lengths, void* stream);

/* ---- span variant TVNetSAModel: models/bert_model.py:113-190, 323-376 ------------------------- */
/* workspace (int32 [2B+1]) := sentence lengths, exclusive prefix sums, total tokens of the compacted stream
 * (get_span_representation :149-152; attention masks are left-aligned). B <= 1024. */
int mtvaf_span_offsets(const int64_t* attention_mask, int B, int L, int32_t* workspace, void* stream);
/* get_span_representation + unary_affine + get_self_att_representation (:147-181, :364-369):
 * pooled[s, :] = sum_j softmax_j(row_j . w + b) row_j over the span's rows of the compacted token stream
 * (span s = (n, m): stream elements offset[n] + start .. + end, clamped to the last element).  seq: [B*L, H] fp32. */
int mtvaf_span_pool_fwd(const float* seq, const int32_t* workspace, const int64_t* span_starts,
                        const int64_t* span_ends, const float* w_unary, const float* b_unary, int B, int L, int M,
                        int H, float* pooled, void* stream);
/* backward: d_seq (fp32 [B*L, H]), d_w_unary [H], d_b_unary [1] are ACCUMULATED (atomics). */
int mtvaf_span_pool_bwd(const float* d_pooled, const float* seq, const int32_t* workspace, const int64_t* span_starts,
                        const int64_t* span_ends, const float* w_unary, const float* b_unary, int B, int L, int M,
                        int H, float* d_seq, float* d_w_unary, float* d_b_unary, void* stream);
/* distant_cross_entropy (:183-192, no mask): loss[0] += scale * mean_b(-sum_l pos log_softmax(x)_l / sum_l pos);
 * logits / dlogits are read / written with an element `stride` (start and end logits are the two columns of the
 * binary_affine output [B*L, 2], :351-354); dlogits (optional) = d(scale * loss)/d logits. */
int mtvaf_distant_ce_fwd_bwd(const float* logits, int64_t stride, const int64_t* positions, int B, int L, float scale,
                             float* loss, float* dlogits, void* stream);
/* nn.CrossEntropyLoss() (mean) over N rows of C classes (:296,302): loss[0] += scale * mean; dlogits optional. */
int mtvaf_ce_mean_fwd_bwd(const float* logits, const int64_t* labels, int N, int C, float scale, float* loss,
                          float* dlogits, void* stream);

/* ---- loss combination: probes/loss.py:13-18 + bert_model.py:523-525 without the .item() sync --- */
/* out[0] = crf_nll_sum/B + (prob_loss > 0.1 ? prob_loss * beta * 2^-epoch : 0) + alpha * img_loss;
 * flag_out[0] = (prob_loss > 0.1).  All scalars live on the device. */
int mtvaf_combine_loss(const float* crf_nll_sum, int B, const float* prob_loss, float beta, int epoch,
                       const float* img_losses, int n_img_losses, float alpha, float* out, int32_t* flag_out,
                       void* stream);

/* ---- optimizer: torch.optim.AdamW as configured in modules/train.py:887-926 ------------------- */
/* zero_grad != 0 clears `grad` in the same pass (replaces optimizer.zero_grad()); bf16_copy (optional) receives
 * the updated weights rounded to bf16 (the tensor-core GEMM operands).  `dyn` (optional, device memory, 24 bytes:
 * uint64 t; float lr_scale, bc1, bc2_sqrt -- maintained by mtvaf_adam_dyn_advance) overrides the by-value step:
 * lr is multiplied by lr_scale and the bias corrections come from the device, so the call can be replayed from
 * a CUDA graph. */
int mtvaf_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                     void* bf16_copy, int zero_grad, const void* dyn, void* stream);
/* dyn->t += 1, then lr_scale = get_linear_schedule_with_warmup (modules/train.py:118-120) for that step
 * (1.0 if total_steps <= 0) and the Adam bias corrections for t. */
int mtvaf_adam_dyn_advance(void* dyn, float beta1, float beta2, int warmup_steps, int total_steps, void* stream);

/* ---- CUDA-graph replay support -------------------------------------------------------------- */
/* Registers a device-resident uint64 step counter (NULL to unregister).  Every dropout site of the library mixes
 * it into its by-value seed, so a training step captured ONCE into a CUDA graph draws fresh dropout masks on every
 * replay (forward and backward of one replay see the same value).  Host-side only: stores the pointer. */
int mtvaf_set_step_source(const uint64_t* dev_step);
/* *dev_step += 1 (stream-ordered; capture it at the head of the graph). */
int mtvaf_advance_step(uint64_t* dev_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTVAF_B200_H_ */
