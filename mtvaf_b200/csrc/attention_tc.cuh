// Shared declarations of the tcgen05 prefix-attention kernels (forward: attention_tc.cu, backward:
// attention_tc_bwd.cu) and of the SIMT implementation they fall back to for shapes they do not cover.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/mtvaf_b200.h"

namespace mtvaf {

struct AttnTcArgs {
  int P, P8;            // prefix rows, padded to a multiple of 8 in the shared-memory key numbering
  int L, L64;           // text rows (queries == keys of the whole item); L64 = text KEY rows loaded, padded to 64
  int N16;              // keys covered by the MMAs: round16(P8 + Lk)
  int Lk, kt0, kbase;   // forward key WINDOW: Lk text keys starting at text row kt0; kbase = reference key number of
                        // the window's first text key (P_total + kt0, dropout hash).  Whole item: Lk = L, kt0 = 0, kbase = P
  int B, nh;
  const long long* key_mask;
  float scale;
  uint32_t drop_thr; float drop_scale; unsigned long long seed;
  const unsigned long long* step;   // device step counter mixed into the seed (or NULL)
};

struct AttnTcMaps {
  CUtensorMap q, kv, kp, vp, dq, dkv;   // q/dq: 128-row boxes, kv/dkv: 64-row boxes, kp/vp: 8-row boxes
};

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer);

int attn_tc_prepare(const void* qkv, int64_t ld_qkv, const void* kp, const void* vp, int P, const int64_t* key_mask,
                    int B, int L, int nh, float p_drop, uint64_t seed, AttnTcArgs* a, AttnTcMaps* m, bool* ok);
int attn_fwd_tc_launch(const AttnTcArgs& a, const AttnTcMaps& m, void* ctx, int64_t ld_ctx, float* lse,
                       cudaStream_t st);
bool attn_fwd_tc_fits(const AttnTcArgs& a);        // all keys of an item resident (N16 <= 448 and shared memory)
// long text (P8 + L > what fits): two key windows through the same kernel + a merge of the partial softmaxes
bool attn_fwd_tc_windows_supported(const AttnTcArgs& a);
size_t attn_fwd_tc_windows_workspace(int B, int L, int nh);
int attn_fwd_tc_windows_launch(const AttnTcArgs& a, const AttnTcMaps& m, void* ctx, int64_t ld_ctx, float* lse,
                               void* workspace, cudaStream_t st);
bool attn_bwd_tc_supported(const AttnTcArgs& a);
int attn_bwd_tc_launch(const AttnTcArgs& a, const AttnTcMaps& m, const void* dctx, int64_t ld_dctx, const void* ctx,
                       int64_t ld_ctx, const float* lse, void* dqkv, int64_t ld_dqkv, float* dkp, float* dvp,
                       cudaStream_t st);
// software-pipelined backward for the reference shape (L <= 128, P <= 16): attention_tc_bwd_pipe.cu
bool attn_bwd_pipe_supported(const AttnTcArgs& a);
// dbias (optional): fp32 [3 * nh * 64], += column sums of dqkv (bias gradient of the fused QKV projection)
int attn_bwd_pipe_launch(const AttnTcArgs& a, const AttnTcMaps& m, const void* dctx, int64_t ld_dctx, const void* ctx,
                         int64_t ld_ctx, const float* lse, void* dqkv, int64_t ld_dqkv, float* dkp, float* dvp,
                         float* dbias, cudaStream_t st);
// short text (L <= 64, P <= 16): two (batch, head) items per tile, attention_tc_bwd_pair.cu
bool attn_bwd_pair_supported(const AttnTcArgs& a);
int attn_bwd_pair_launch(const AttnTcArgs& a, const AttnTcMaps& m, const void* dctx, int64_t ld_dctx, const void* ctx,
                         int64_t ld_ctx, const float* lse, void* dqkv, int64_t ld_dqkv, float* dkp, float* dvp,
                         float* dbias, cudaStream_t st);
// long-text backward (128 < L <= 512), attention_tc_bwd_long.cu
bool attn_bwd_long_supported(const AttnTcArgs& a);
int attn_bwd_long_launch(const AttnTcArgs& a, const AttnTcMaps& m, const void* dctx, int64_t ld_dctx, const void* ctx,
                         int64_t ld_ctx, const float* lse, void* dqkv, int64_t ld_dqkv, float* dkp, float* dvp,
                         cudaStream_t st);
int attention_impl_override();   // 0 = auto (tcgen05 when the shape fits), 1 = SIMT only, 2 = tcgen05 without the pipelined bwd

}  // namespace mtvaf
