"""Thin torch-tensor wrappers over the C ABI (include/mtvaf_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every arithmetic op below is one of the
library's hand-written sm_100a kernels.  No wrapper has a torch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L

F32, BF16 = L.F32, L.BF16
_raw = L.raw()


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError("mtvaf_b200 supports float32 and bfloat16 activations, got %s" % t.dtype)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


# the library counts its own kernel launches (mtvaf_launch_count); the bench reports the count of OUR launches
_launch_base = 0
GEMM_EVENT_SINK = None      # bench.py: list receiving (start_event, end_event, flops, algorithmic bytes, (M, N, K, a_mn, b_mn, mode)) per tcgen05 GEMM launch


def reset_launch_count():
    global _launch_base
    _launch_base = int(_raw.mtvaf_launch_count())


def launch_count() -> int:
    return int(_raw.mtvaf_launch_count()) - _launch_base


def _check(rc: int, name: str):
    if rc != 0:
        raise L.MtvafError("%s failed (%d): %s" % (name, rc, L.last_error()))


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.MtvafError("mtvaf_b200 ops need CUDA tensors (no CPU fallback exists)")


# ------------------------------------------------------------------------------------------ GEMM
def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, M: int, N: int, K: int,
         mode: int = L.EPI_STORE, out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
         bias: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
         out2: Optional[torch.Tensor] = None, rowvec: Optional[torch.Tensor] = None, alpha: float = 1.0,
         p_drop: float = 0.0, seed: int = 0, splits: int = 1, ldo: Optional[int] = None,
         colsum: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """D[M,N] = epilogue(alpha * A B^T) -- see MtvafEpilogue in the header. a, b are 2-D row-major
    tensors: a is [M,K] (a_mn=False) or [K,M] (a_mn=True); same for b with N."""
    _cuda(a, b, out, bias, aux, out2, rowvec)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    assert a.dtype == b.dtype
    if out is None and mode != L.EPI_SQNORM:
        od = out_dtype or (torch.float32 if mode == L.EPI_ATOMIC_F32 else a.dtype)
        out = torch.empty((M, N), dtype=od, device=a.device)
    ep = L.Epilogue()
    ep.mode = mode
    ep.out_dtype = dt(out) if out is not None else F32
    ep.out = _p(out)
    ep.ldo = (ldo if ldo is not None else out.stride(0)) if out is not None else 0
    ep.bias = _p(bias)
    ep.aux = _p(aux)
    ep.ld_aux = aux.stride(0) if aux is not None else 0
    ep.out2 = _p(out2)
    ep.ld_out2 = out2.stride(0) if out2 is not None else 0
    ep.rowvec = _p(rowvec)
    ep.alpha = alpha
    ep.p_drop = p_drop
    ep.seed = seed
    ep.colsum = _p(colsum)
    if colsum is not None:
        assert colsum.dtype == torch.float32 and colsum.is_contiguous() and colsum.numel() >= N
    if bias is not None:
        assert bias.dtype == torch.float32
    if aux is not None:
        assert aux.dtype == a.dtype
    fn = _raw.mtvaf_gemm_bf16 if a.dtype == torch.bfloat16 else _raw.mtvaf_gemm_f32
    sink = GEMM_EVENT_SINK if a.dtype == torch.bfloat16 else None
    if sink is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = fn(a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn), M, N, K, C.byref(ep),
            splits, _stream())
    if sink is not None:
        e1.record()
        nbytes = 2 * (M * K + N * K) + (M * N * out.element_size() if out is not None else 0)
        nbytes += (M * N * 2 if aux is not None else 0) + (M * N * 2 if out2 is not None else 0)
        sink.append((e0, e1, 2.0 * M * N * K, nbytes, (M, N, K, int(a_mn), int(b_mn), int(mode))))
    _check(rc, "mtvaf_gemm")
    return out


def set_gemm_impl(impl: str):
    """'auto' (CTA-pair cta_group::2 kernel when M >= 256) or 'single' (single-CTA tiles; A/B testing)."""
    global _GEMM_PAIR
    _check(_raw.mtvaf_set_gemm_impl({"auto": 0, "single": 1}[impl]), "set_gemm_impl")
    _GEMM_PAIR = impl == "auto"


_GEMM_PAIR = True


_SM_RESERVE = 0


def set_sm_reserve(n_sms: int):
    """SMs the persistent kernels leave to a concurrent collective (whole TPCs; 0 = none)."""
    global _SM_RESERVE
    _check(_raw.mtvaf_set_sm_reserve(int(n_sms)), "set_sm_reserve")
    _SM_RESERVE = int(n_sms) & ~1


def linear_fwd(x, w, bias=None, **kw):
    """y = x w^T + b  (x [M,K], w [N,K])"""
    return gemm(x, w, M=x.shape[0], N=w.shape[0], K=x.shape[1], bias=bias, **kw)


def linear_dgrad(dy, w, **kw):
    """dx = dy w  (dy [M,N], w [N,K])"""
    return gemm(dy, w, b_mn=True, M=dy.shape[0], N=w.shape[1], K=dy.shape[1], **kw)


def wgrad_splits(m_out: int, n_out: int, k: int, bf16: bool) -> int:
    bm, bn = (128, 256 if n_out > 128 else 128) if bf16 else (128, 128)
    sms = 148 - _SM_RESERVE
    slots = sms * (1 if bf16 else 2)          # persistent CTAs (bf16) / resident blocks (fp32)
    if bf16 and _GEMM_PAIR and m_out >= 256:
        bm, slots = 256, sms // 2             # CTA-pair kernel: 256-row tiles, one cluster per TPC
    tiles = ((m_out + bm - 1) // bm) * ((n_out + bn - 1) // bn)
    kb = max(1, k // (64 if bf16 else 16))
    best, best_score = 1, -1.0
    for s in range(1, min(kb, 32) + 1):
        items = tiles * s
        waves = (items + slots - 1) // slots
        score = items / float(waves * slots) - 0.01 * s      # wave efficiency, mild penalty for atomics traffic
        if score > best_score:
            best, best_score = s, score
    return best


def skinny_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 y = x w^T + b for N = w.shape[0] <= 48: deterministic warp-per-row kernel (no split-K atomics)."""
    _cuda(x, w, bias)
    assert x.dtype == torch.float32 and w.dtype == torch.float32 and x.stride(1) == 1 and w.stride(1) == 1
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    _check(_raw.mtvaf_skinny_linear_f32(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _p(bias), M, N, K,
                                        out.data_ptr(), out.stride(0), _stream()), "skinny_linear")
    return out


def skinny_linear_dgrad(dy: torch.Tensor, w: torch.Tensor, out_dtype: torch.dtype, p_drop: float = 0.0,
                        seed: int = 0, tanh_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx = dropout_mask(seed) * (dy w) / (1-p) [* (1 - tanh_out^2)] for N = w.shape[0] <= 16, in `out_dtype`."""
    _cuda(dy, w, tanh_out)
    assert dy.dtype == torch.float32 and w.dtype == torch.float32 and dy.stride(1) == 1 and w.stride(1) == 1
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty((M, K), dtype=out_dtype, device=dy.device)
    if tanh_out is not None:
        assert tanh_out.dtype == out_dtype and tanh_out.shape == dx.shape and tanh_out.is_contiguous()
    _check(_raw.mtvaf_skinny_linear_dgrad(dy.data_ptr(), dy.stride(0), w.data_ptr(), w.stride(0), M, N, K, p_drop,
                                          seed, dx.data_ptr(), dx.stride(0), dt(dx), _p(tanh_out), _stream()),
           "skinny_linear_dgrad")
    return dx


def skinny_linear_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor):
    """dw[N,K] += dy^T x (fp32) for skinny N."""
    _cuda(dy, x, dw)
    assert dy.dtype == torch.float32 and x.dtype == torch.float32 and dw.dtype == torch.float32
    M, N = dy.shape
    K = x.shape[1]
    _check(_raw.mtvaf_skinny_linear_wgrad(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), M, N, K,
                                          dw.data_ptr(), dw.stride(0), _stream()), "skinny_linear_wgrad")


def skinny_splits(M: int, N: int, K: int) -> int:
    """split-K factor for the fp32 SIMT GEMM when the output has too few 128x128 tiles to fill the SMs."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    return max(1, min(K // 64, (2 * 148) // max(1, tiles)))


def linear_wgrad(dy, x, dw: torch.Tensor, *, n_valid: Optional[int] = None):
    """dw[N,K] += dy^T x  (dy [M,N], x [M,K]); fp32 atomics into the gradient buffer."""
    M = dy.shape[0]
    N = n_valid if n_valid is not None else dy.shape[1]
    K = x.shape[1]
    gemm(dy, x, a_mn=True, b_mn=True, M=N, N=K, K=M, mode=L.EPI_ATOMIC_F32, out=dw,
         splits=wgrad_splits(N, K, M, dy.dtype == torch.bfloat16))


def colsum(dy: torch.Tensor, db: torch.Tensor, n_valid: Optional[int] = None):
    _cuda(dy, db)
    N = n_valid if n_valid is not None else dy.shape[1]
    _check(_raw.mtvaf_colsum(dy.data_ptr(), dy.stride(0), dt(dy), dy.shape[0], N, db.data_ptr(), _stream()), "colsum")


def dropout_apply(x: torch.Tensor, p_drop: float, seed: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(x)
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    _check(_raw.mtvaf_dropout_apply(x.data_ptr(), out.data_ptr(), x.numel(), dt(x), p_drop, seed, _stream()),
           "dropout_apply")
    return out


def add_inplace(dst: torch.Tensor, src: torch.Tensor, alpha: float = 1.0):
    _cuda(dst, src)
    assert dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()
    _check(_raw.mtvaf_add_inplace(dst.data_ptr(), dt(dst), src.data_ptr(), dt(src), dst.numel(), alpha, _stream()),
           "add_inplace")


def rowscale(x: torch.Tensor, rs: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    _cuda(x, rs)
    assert x.is_contiguous() and rs.dtype == torch.float32
    y = torch.empty_like(x)
    _check(_raw.mtvaf_rowscale(x.data_ptr(), rs.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], alpha, dt(x),
                               _stream()), "rowscale")
    return y


def scale_by_device_scalar(x: torch.Tensor, s: torch.Tensor):
    _cuda(x, s)
    assert x.dtype == torch.float32 and s.dtype == torch.float32 and x.is_contiguous()
    _check(_raw.mtvaf_scale_by_device_scalar(x.data_ptr(), x.numel(), s.data_ptr(), _stream()), "scale_dev")


def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    _check(_raw.mtvaf_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "cast")
    return dst


def cast_f32(src: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(src)
    assert src.dtype == torch.bfloat16 and src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    _check(_raw.mtvaf_cast_bf16_to_f32(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "cast")
    return dst


# ------------------------------------------------------------------------------------------ LN / embeddings
def layernorm_fwd(z, gamma, beta, eps, y=None):
    _cuda(z, gamma, beta)
    rows, H = z.shape
    if y is None:
        y = torch.empty_like(z)
    mean = torch.empty(rows, dtype=torch.float32, device=z.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=z.device)
    _check(_raw.mtvaf_layernorm_fwd(z.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, rows, H,
                                    dt(z), mean.data_ptr(), rstd.data_ptr(), _stream()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, z, gamma, mean, rstd, d_gamma, d_beta, dz=None, d_bias=None, p_drop=0.0, seed=0):
    """Returns (dz, dd): dd = dropout-masked dz (the gradient entering the preceding dense layer; dd is dz when
    p_drop == 0); d_bias (optional) += column sums of dd."""
    rows, H = z.shape
    if dz is None:
        dz = torch.empty_like(z)
    dd = torch.empty_like(z) if p_drop > 0 else None
    _check(_raw.mtvaf_layernorm_bwd(dy.data_ptr(), z.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    rows, H, dt(z), dz.data_ptr(), d_gamma.data_ptr(), d_beta.data_ptr(), _p(dd),
                                    _p(d_bias), p_drop, seed, _stream()), "layernorm_bwd")
    return dz, (dd if dd is not None else dz)


def embed_ln_fwd(ids, tts, word, pos, typ, gamma, beta, eps, kind, pad_idx, out_dtype, p_drop=0.0, seed=0):
    _cuda(ids, tts, word)
    B, Lq = ids.shape
    H = word.shape[1]
    out = torch.empty((B * Lq, H), dtype=out_dtype, device=ids.device)
    pids = torch.empty((B, Lq), dtype=torch.int64, device=ids.device)
    mean = torch.empty(B * Lq, dtype=torch.float32, device=ids.device)
    rstd = torch.empty(B * Lq, dtype=torch.float32, device=ids.device)
    _check(_raw.mtvaf_embed_ln_fwd(ids.data_ptr(), tts.data_ptr(), word.data_ptr(), pos.data_ptr(), typ.data_ptr(),
                                   gamma.data_ptr(), beta.data_ptr(), eps, kind, pad_idx, B, Lq, H, word.shape[0],
                                   pos.shape[0], typ.shape[0], out.data_ptr(), dt(out), pids.data_ptr(),
                                   mean.data_ptr(), rstd.data_ptr(), p_drop, seed, _stream()), "embed_ln_fwd")
    return out, pids, mean, rstd


def embed_ln_bwd(dout, ids, tts, pids, word, pos, typ, gamma, mean, rstd, kind, pad_idx, d_word, d_pos, d_type,
                 d_gamma, d_beta, p_drop=0.0, seed=0):
    B, Lq = ids.shape
    H = word.shape[1]
    _check(_raw.mtvaf_embed_ln_bwd(dout.data_ptr(), dt(dout), ids.data_ptr(), tts.data_ptr(), pids.data_ptr(),
                                   word.data_ptr(), pos.data_ptr(), typ.data_ptr(), gamma.data_ptr(), mean.data_ptr(),
                                   rstd.data_ptr(), kind, pad_idx, B, Lq, H, d_word.data_ptr(), d_pos.data_ptr(),
                                   d_type.data_ptr(), d_gamma.data_ptr(), d_beta.data_ptr(), p_drop, seed, _stream()),
           "embed_ln_bwd")


# ------------------------------------------------------------------------------------------ attention
def set_attention_impl(impl: str):
    """'auto' (tcgen05 kernels when bf16 and the shape fits), 'simt', or 'tc_generic' (tcgen05 kernels with the
    generic instead of the software-pipelined backward) -- A/B testing."""
    _check(_raw.mtvaf_set_attention_impl({"auto": 0, "simt": 1, "tc_generic": 2}[impl]), "set_attention_impl")


def attention_fwd(qkv, kp, vp, key_mask, B, Lq, nh, d, p_drop=0.0, seed=0, want_probs=False):
    _cuda(qkv, key_mask)
    P = 0 if kp is None else kp.shape[2]
    ctx = torch.empty((B * Lq, nh * d), dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty((B, nh, Lq), dtype=torch.float32, device=qkv.device)
    probs = torch.empty((B, nh, Lq, P + Lq), dtype=torch.float32, device=qkv.device) if want_probs else None
    # long text (keys of an item beyond one resident tile set): caller-owned workspace for the two-window forward
    ws_bytes = int(_raw.mtvaf_attention_fwd_workspace_bytes(B, Lq, nh, d, P, dt(qkv)))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=qkv.device) if ws_bytes > 0 else None
    _check(_raw.mtvaf_attention_fwd_ws(qkv.data_ptr(), qkv.stride(0), _p(kp), _p(vp), P, key_mask.data_ptr(), B, Lq,
                                       nh, d, ctx.data_ptr(), ctx.stride(0), lse.data_ptr(), _p(probs), dt(qkv),
                                       p_drop, seed, _p(ws), ws_bytes, _stream()), "attention_fwd")
    return ctx, lse, probs


def attention_bwd(dctx, qkv, kp, vp, key_mask, ctx, lse, B, Lq, nh, d, dkp=None, dvp=None, p_drop=0.0, seed=0,
                  dqkv=None, d_bias=None):
    """d_bias (optional): fp32 [3*nh*d], += column sums of dqkv (bias gradient of the fused QKV projection)."""
    P = 0 if kp is None else kp.shape[2]
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    scratch = torch.empty((B, nh, Lq), dtype=torch.float32, device=qkv.device)
    if d_bias is not None:
        assert d_bias.dtype == torch.float32 and d_bias.numel() == 3 * nh * d and d_bias.is_contiguous()
    _check(_raw.mtvaf_attention_bwd_ex(dctx.data_ptr(), dctx.stride(0), qkv.data_ptr(), qkv.stride(0), _p(kp), _p(vp),
                                       P, key_mask.data_ptr(), ctx.data_ptr(), ctx.stride(0), lse.data_ptr(), B, Lq,
                                       nh, d, dqkv.data_ptr(), dqkv.stride(0), _p(dkp), _p(dvp), scratch.data_ptr(),
                                       dt(qkv), p_drop, seed, _p(d_bias), _stream()), "attention_bwd")
    return dqkv


# ------------------------------------------------------------------------------------------ fusion / heads
def mean4_fwd(x, rows, W, mode):
    y = torch.empty((rows, W), dtype=x.dtype, device=x.device)
    _check(_raw.mtvaf_mean4_fwd(x.data_ptr(), y.data_ptr(), rows, W, mode, dt(x), _stream()), "mean4_fwd")
    return y


def mean4_bwd_add(dy, dx, rows, W, mode):
    assert dy.dtype == torch.float32 and dx.dtype == torch.float32
    _check(_raw.mtvaf_mean4_bwd_add(dy.data_ptr(), dx.data_ptr(), rows, W, mode, _stream()), "mean4_bwd")


def gate_fwd(guids, gate_logits, n_layers, n_img, B, hid):
    P = 4 * n_img
    kv = torch.empty((n_layers, 2, B, P * hid), dtype=guids.dtype, device=guids.device)
    gates = torch.empty((n_img * B, n_layers * 4), dtype=torch.float32, device=guids.device)
    _check(_raw.mtvaf_gate_fwd(guids.data_ptr(), gate_logits.data_ptr(), n_layers, n_img, B, hid, kv.data_ptr(),
                               gates.data_ptr(), dt(guids), _stream()), "gate_fwd")
    return kv, gates


def gate_bwd(d_kv, guids, gate_logits, gates, n_layers, n_img, B, hid, d_guids):
    scratch = torch.zeros_like(gates)
    d_logits = torch.empty_like(gates)
    _check(_raw.mtvaf_gate_bwd(d_kv.data_ptr(), guids.data_ptr(), gate_logits.data_ptr(), gates.data_ptr(), n_layers,
                               n_img, B, hid, d_guids.data_ptr(), scratch.data_ptr(), d_logits.data_ptr(), dt(guids),
                               _stream()), "gate_bwd")
    return d_logits


def prompt_grad_combine(d_guids, d_gs, d_gm, p_drop, seed, rows, W, out_dtype):
    """d(prompt) = gate path + both 4-way-mean backward terms, one pass, written in `out_dtype`."""
    out = torch.empty((rows * 4, W), dtype=out_dtype, device=d_guids.device)
    _check(_raw.mtvaf_prompt_grad_combine(d_guids.data_ptr(), _p(d_gs), _p(d_gm), dt(d_gm) if d_gm is not None else F32,
                                          p_drop, seed, rows, W, out.data_ptr(), dt(out), _stream()),
           "prompt_grad_combine")
    return out


def softmax_kl(logits, n, target, B, want_grad, grad_scale=1.0):
    rows = logits.shape[0]
    loss = torch.zeros(rows // B, dtype=torch.float32, device=logits.device)
    dlogits = torch.zeros_like(logits) if want_grad else None
    _check(_raw.mtvaf_softmax_kl_fwd_bwd(logits.data_ptr(), logits.stride(0), target.data_ptr(), rows, B, n,
                                         loss.data_ptr(), _p(dlogits), grad_scale, _stream()), "softmax_kl")
    return loss, dlogits


def row_sqnorm(x: torch.Tensor) -> torch.Tensor:
    """out[r] = sum_c x[r, c]^2 (fp32) for a 2-D bf16 / fp32 tensor with cols % 8 == 0."""
    _cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _check(_raw.mtvaf_row_sqnorm(x.data_ptr(), x.stride(0), dt(x), x.shape[0], x.shape[1], out.data_ptr(), _stream()),
           "row_sqnorm")
    return out


def probe_labels(norms):
    B, Lq = norms.shape
    labels = torch.empty_like(norms)
    _check(_raw.mtvaf_probe_labels(norms.data_ptr(), labels.data_ptr(), B, Lq, _stream()), "probe_labels")
    return labels


def mse(norms, labels, want_grad):
    loss = torch.empty(1, dtype=torch.float32, device=norms.device)
    dn = torch.empty_like(norms) if want_grad else None
    _check(_raw.mtvaf_mse_fwd_bwd(norms.data_ptr(), labels.data_ptr(), norms.numel(), loss.data_ptr(), _p(dn),
                                  _stream()), "mse")
    return loss, dn


def pairwise_sqdist(T, B, Lq, R):
    dist = torch.empty((B, Lq, Lq), dtype=torch.float32, device=T.device)
    _check(_raw.mtvaf_pairwise_sqdist(T.data_ptr(), T.stride(0), dt(T), B, Lq, R, dist.data_ptr(), _stream()),
           "pairwise_sqdist")
    return dist


def set_pairwise_impl(impl: str):
    """'auto' (tcgen05 Gram form for fp32 T with R % 64 == 0) or 'simt' (explicit differences) -- A/B testing."""
    _check(_raw.mtvaf_set_pairwise_impl({"auto": 0, "simt": 1}[impl]), "set_pairwise_impl")


def pack_features(images: torch.Tensor, aux_imgs: Optional[torch.Tensor], out_dtype: torch.dtype) -> torch.Tensor:
    """Feature wire format -> GEMM operand: images [B, E...] and aux_imgs [B, n_aux, E...] (fp32 or bf16, E = 3840*2*2
    elements per image) -> [1 + n_aux, B, E] in `out_dtype`; row j*B + b = image j of sample b (models/bert_model.py:
    536-539: cat of the pyramid levels + plain view, aux images permuted to the front)."""
    _cuda(images, aux_imgs)
    B = images.shape[0]
    E = images.numel() // B
    n_aux = 0 if aux_imgs is None else aux_imgs.shape[1]
    # each sample's block must be dense; the stride BETWEEN samples is free (views of one [B, 1+n_aux, E] wire buffer)
    assert images[0].is_contiguous()
    assert aux_imgs is None or (aux_imgs[0].is_contiguous() and aux_imgs.dtype == images.dtype
                                and aux_imgs.numel() == B * n_aux * E)
    out = torch.empty((1 + n_aux, B, E), dtype=out_dtype, device=images.device)
    _check(_raw.mtvaf_pack_features(images.data_ptr(), images.stride(0), _p(aux_imgs),
                                    0 if aux_imgs is None else aux_imgs.stride(0), dt(images), B, n_aux, E,
                                    out.data_ptr(), dt(out), _stream()), "pack_features")
    return out


def crf_nll(em, tags, mask, start, end, trans, want_grad, grad_scale):
    B, Lq, T = em.shape
    nll = torch.zeros(1, dtype=torch.float32, device=em.device)
    if want_grad:
        d_em = torch.empty_like(em)
        d_s, d_e, d_t = torch.zeros_like(start), torch.zeros_like(end), torch.zeros_like(trans)
    else:
        d_em = d_s = d_e = d_t = None
    _check(_raw.mtvaf_crf_nll_fwd_bwd(em.data_ptr(), tags.data_ptr(), mask.data_ptr(), start.data_ptr(),
                                      end.data_ptr(), trans.data_ptr(), B, Lq, T, nll.data_ptr(), _p(d_em), _p(d_s),
                                      _p(d_e), _p(d_t), grad_scale, _stream()), "crf_nll")
    return nll, d_em, d_s, d_e, d_t


def crf_decode(em, mask, start, end, trans):
    B, Lq, T = em.shape
    best = torch.empty((B, Lq), dtype=torch.int64, device=em.device)
    lens = torch.empty(B, dtype=torch.int64, device=em.device)
    _check(_raw.mtvaf_crf_decode(em.data_ptr(), mask.data_ptr(), start.data_ptr(), end.data_ptr(), trans.data_ptr(),
                                 B, Lq, T, best.data_ptr(), lens.data_ptr(), _stream()), "crf_decode")
    return best, lens


# ------------------------------------------------------------------------------------------ span variant heads
def span_offsets(mask):
    B, Lq = mask.shape
    ws = torch.empty(2 * B + 1, dtype=torch.int32, device=mask.device)
    _check(_raw.mtvaf_span_offsets(mask.data_ptr(), B, Lq, ws.data_ptr(), _stream()), "span_offsets")
    return ws


def span_pool_fwd(seq, ws, starts, ends, w_u, b_u, B, Lq):
    M = starts.shape[1]
    H = seq.shape[1]
    assert seq.dtype == torch.float32 and seq.is_contiguous()
    pooled = torch.empty((B * M, H), dtype=torch.float32, device=seq.device)
    _check(_raw.mtvaf_span_pool_fwd(seq.data_ptr(), ws.data_ptr(), starts.data_ptr(), ends.data_ptr(), w_u.data_ptr(),
                                    b_u.data_ptr(), B, Lq, M, H, pooled.data_ptr(), _stream()), "span_pool_fwd")
    return pooled


def span_pool_bwd(d_pooled, seq, ws, starts, ends, w_u, b_u, B, Lq, d_seq, d_w, d_b):
    M = starts.shape[1]
    H = seq.shape[1]
    _check(_raw.mtvaf_span_pool_bwd(d_pooled.data_ptr(), seq.data_ptr(), ws.data_ptr(), starts.data_ptr(),
                                    ends.data_ptr(), w_u.data_ptr(), b_u.data_ptr(), B, Lq, M, H, d_seq.data_ptr(),
                                    d_w.data_ptr(), d_b.data_ptr(), _stream()), "span_pool_bwd")


def distant_ce(logits2, col, positions, scale, loss, dlogits2):
    """logits2: [B*L, 2] fp32 (binary_affine output); col 0 = start, 1 = end. Accumulates into loss[0]."""
    B, Lq = positions.shape
    d = None if dlogits2 is None else dlogits2.data_ptr() + 4 * col
    _check(_raw.mtvaf_distant_ce_fwd_bwd(logits2.data_ptr() + 4 * col, logits2.stride(0), positions.data_ptr(), B, Lq,
                                         scale, loss.data_ptr(), d, _stream()), "distant_ce")


def ce_mean(logits, labels, scale, loss, want_grad):
    N, Cc = logits.shape
    d = torch.empty_like(logits) if want_grad else None
    _check(_raw.mtvaf_ce_mean_fwd_bwd(logits.data_ptr(), labels.data_ptr(), N, Cc, scale, loss.data_ptr(), _p(d),
                                      _stream()), "ce_mean")
    return d


def combine_loss(crf_nll_sum, B, prob_loss, beta, epoch, img_losses, alpha):
    out = torch.empty(1, dtype=torch.float32, device=crf_nll_sum.device)
    flag = torch.empty(1, dtype=torch.int32, device=crf_nll_sum.device)
    n_img = 0 if img_losses is None else img_losses.numel()
    _check(_raw.mtvaf_combine_loss(crf_nll_sum.data_ptr(), B, _p(prob_loss), beta, epoch, _p(img_losses), n_img,
                                   alpha, out.data_ptr(), flag.data_ptr(), _stream()), "combine_loss")
    return out, flag


def adamw_step(param, grad, m, v, lr, b1, b2, eps, wd, step, grad_scale=1.0, bf16_copy=None, zero_grad=False,
               dyn=None):
    _check(_raw.mtvaf_adamw_step(param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), param.numel(), lr, b1,
                                 b2, eps, wd, step, grad_scale, _p(bf16_copy), int(zero_grad), _p(dyn), _stream()),
           "adamw_step")


def adam_dyn_advance(dyn, b1, b2, warmup_steps, total_steps):
    """dyn: device int64[3] buffer (24 bytes) holding {uint64 t; float lr_scale, bc1, bc2_sqrt}."""
    _check(_raw.mtvaf_adam_dyn_advance(dyn.data_ptr(), b1, b2, int(warmup_steps), int(total_steps), _stream()),
           "adam_dyn_advance")


_STEP_SOURCE_PTR = 0


def set_step_source(t: Optional[torch.Tensor]):
    """Register (or clear) the device step counter every dropout site mixes into its seed (CUDA-graph replay).
    The registration is process-global (one training step per process)."""
    global _STEP_SOURCE_PTR
    if t is not None:
        assert t.is_cuda and t.dtype == torch.int64 and t.numel() == 1
    _check(_raw.mtvaf_set_step_source(_p(t)), "set_step_source")
    _STEP_SOURCE_PTR = 0 if t is None else t.data_ptr()


def clear_step_source_if(ptr: int):
    """Unregister the step counter iff `ptr` is still the registered one (finalizer of GraphedTrainStep: a dropped
    step object must not leave a dangling device pointer behind, nor clear a newer object's registration)."""
    if ptr and ptr == _STEP_SOURCE_PTR:
        set_step_source(None)


def advance_step(t: torch.Tensor):
    _check(_raw.mtvaf_advance_step(t.data_ptr(), _stream()), "advance_step")
