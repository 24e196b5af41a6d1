"""GPU parity of every C-ABI kernel against the oracle / plain torch fp32 on the same seeded inputs.
Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import math

import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import mtvaf_oracle as O   # checker only


@pytest.fixture(scope="module")
def ops():
    from mtvaf_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def Lb():
    from mtvaf_b200 import lib
    return lib


DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ---------------------------------------------------------------------------------------- GEMM
GEMM_SHAPES = [(128, 256, 64), (256, 768, 768), (384, 2304, 768), (200, 800, 3840), (64, 6144, 800),
               (130, 48, 6144), (512, 384, 768), (96, 2089, 6144), (2048, 768, 3072)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_forward_layout(ops, Lb, dtype, M, N, K):
    x = rnd(M, K, seed=1, dtype=dtype)
    w = rnd(N, K, seed=2, scale=0.05, dtype=dtype)
    bias = rnd(N, seed=3)
    y = ops.linear_fwd(x, w, bias)
    ref = x.float() @ w.float().t() + bias
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-5
    assert rel_err(y, ref) < tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K", [(256, 768, 768), (200, 768, 3072), (384, 3072, 768), (64, 3840, 800), (96, 6144, 2089)])
def test_gemm_dgrad_layout(ops, Lb, dtype, M, N, K):
    # dX[M,N] = dY[M,K] W[K,N]   (B operand MN-major)
    ld = (K + 7) // 8 * 8
    dy_full = torch.zeros(M, ld, dtype=dtype, device=DEV)
    dy_full[:, :K] = rnd(M, K, seed=4, dtype=dtype)
    dy = dy_full[:, :K]
    w = rnd(K, N, seed=5, scale=0.05, dtype=dtype)
    out = ops.gemm(dy, w, b_mn=True, M=M, N=N, K=K)
    ref = dy.float() @ w.float()
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-5
    assert rel_err(out, ref) < tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("T,N,K", [(256, 768, 768), (1000, 3072, 768), (4096, 768, 3072), (64, 800, 3840), (96, 48, 6144)])
def test_gemm_wgrad_layout(ops, Lb, dtype, T, N, K):
    # dW[N,K] += dY[T,N]^T X[T,K]  (both MN-major, split-K fp32 atomics)
    dy = rnd(T, N, seed=6, dtype=dtype)
    x = rnd(T, K, seed=7, dtype=dtype)
    dw = torch.zeros(N, K, dtype=torch.float32, device=DEV)
    ops.linear_wgrad(dy, x, dw)
    ref = dy.float().t() @ x.float()
    tol = 1e-2 if dtype == torch.bfloat16 else 5e-5
    assert rel_err(dw, ref) < tol
    ops.linear_wgrad(dy, x, dw)       # accumulates
    assert rel_err(dw, 2 * ref) < tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K", [(1000, 3072, 768), (300, 768, 256), (130, 200, 64), (4099, 2304, 768)])
def test_gemm_fused_column_sums(ops, Lb, dtype, M, N, K):
    """MtvafEpilogue.colsum: += column sums of the stored output, fused (CTA-pair staged epilogue, M >= 256, bf16) or
    as a post-pass (every other path); partial last row tile and ragged N included."""
    dy = rnd(M, K, seed=31, dtype=dtype)
    w = rnd(K, N, seed=32, scale=0.05, dtype=dtype)
    aux = rnd(M, N, seed=33, dtype=dtype)
    base = rnd(N, seed=34)
    cs = base.clone()
    out = ops.gemm(dy, w, b_mn=True, M=M, N=N, K=K, mode=Lb.EPI_MUL_DGELU, aux=aux, colsum=cs)
    ref = out.float().sum(0)
    assert rel_err(cs - base, ref) < (2e-3 if dtype == torch.bfloat16 else 2e-5)
    out2 = ops.gemm(dy, w, b_mn=True, M=M, N=N, K=K, mode=Lb.EPI_MUL_DGELU, aux=aux)
    assert torch.equal(out, out2)
    cs3 = base.clone()                                                      # the training path's form: x * saved gelu'
    out3 = ops.gemm(dy, w, b_mn=True, M=M, N=N, K=K, mode=Lb.EPI_MUL_AUX, aux=aux, colsum=cs3)
    assert rel_err(out3, (dy.float() @ w.float()) * aux.float()) < (1.5e-2 if dtype == torch.bfloat16 else 3e-5)
    assert rel_err(cs3 - base, out3.float().sum(0)) < (2e-3 if dtype == torch.bfloat16 else 2e-5)
    cs2 = torch.zeros(N, device=DEV)
    bias = rnd(N, seed=35)
    y = ops.linear_fwd(dy, w.t().contiguous(), bias, colsum=cs2)          # STORE + bias: rows past M must not count
    assert rel_err(cs2, y.float().sum(0)) < (2e-3 if dtype == torch.bfloat16 else 2e-5)


def test_gemm_tag_head_shapes_bf16(ops, Lb):
    """The tag head fc 768 -> 11 (models/bert_model.py:510) on the tcgen05 GEMMs in bf16 mode: N = 11 forward with fp32
    emissions, K = 11 data gradient, M = 11 weight gradient from a [T,16] zero-padded d(emissions) operand."""
    T, H, n_tags = 4099, 768, 11
    x = rnd(T, H, seed=11, dtype=torch.bfloat16)
    w = rnd(n_tags, H, seed=12, scale=0.05, dtype=torch.bfloat16)
    bias = rnd(n_tags, seed=13)
    em = ops.linear_fwd(x, w, bias, out_dtype=torch.float32)
    assert em.dtype == torch.float32 and em.shape == (T, n_tags)
    assert rel_err(em, x.float() @ w.float().t() + bias) < 1e-2
    de16 = torch.zeros(T, 16, dtype=torch.bfloat16, device=DEV)
    de16[:, :n_tags] = rnd(T, n_tags, seed=14, dtype=torch.bfloat16)
    dw = torch.zeros(n_tags, H, dtype=torch.float32, device=DEV)
    guard = dw.clone()
    ops.gemm(de16, x, a_mn=True, b_mn=True, M=n_tags, N=H, K=T, mode=Lb.EPI_ATOMIC_F32, out=dw,
             splits=ops.wgrad_splits(n_tags, H, T, True))
    assert rel_err(dw, de16[:, :n_tags].float().t() @ x.float()) < 1e-2
    dx = ops.gemm(de16, w, b_mn=True, M=T, N=H, K=n_tags)
    assert rel_err(dx, de16[:, :n_tags].float() @ w.float()) < 1e-2
    assert torch.equal(guard, torch.zeros_like(guard))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_epilogues(ops, Lb, dtype):
    M, N, K = 300, 768, 256
    tol = 1.5e-2 if dtype == torch.bfloat16 else 3e-5
    x = rnd(M, K, seed=1, dtype=dtype)
    w = rnd(N, K, seed=2, scale=0.08, dtype=dtype)
    bias = rnd(N, seed=3)
    base = x.float() @ w.float().t() + bias
    # GELU with pre-activation side output
    pre = torch.empty(M, N, dtype=dtype, device=DEV)
    g = ops.linear_fwd(x, w, bias, mode=Lb.EPI_GELU, out2=pre)
    assert rel_err(pre, base) < tol
    assert rel_err(g, F.gelu(base)) < tol
    # tanh
    t = ops.linear_fwd(x, w, bias, mode=Lb.EPI_TANH)
    assert rel_err(t, torch.tanh(base)) < tol
    # residual (no dropout)
    res = rnd(M, N, seed=9, dtype=dtype)
    r = ops.linear_fwd(x, w, bias, mode=Lb.EPI_RESID, aux=res)
    assert rel_err(r, base + res.float()) < tol
    # residual with dropout: kept fraction and scaling
    r2 = ops.linear_fwd(x, w, bias, mode=Lb.EPI_RESID, aux=torch.zeros_like(res), p_drop=0.25, seed=1234)
    kept = (r2.float() != 0).float().mean().item()
    assert abs(kept - 0.75) < 0.01
    mask = r2.float() != 0
    assert rel_err(r2.float()[mask], (base / 0.75)[mask]) < tol
    r3 = ops.linear_fwd(x, w, bias, mode=Lb.EPI_RESID, aux=torch.zeros_like(res), p_drop=0.25, seed=1234)
    assert torch.equal(r2, r3)                      # counter-based: reproducible
    # dgelu multiply
    pre_in = rnd(M, N, seed=11, dtype=dtype)
    dg = ops.linear_fwd(x, w, None, mode=Lb.EPI_MUL_DGELU, aux=pre_in)
    pf = pre_in.float().requires_grad_()
    F.gelu(pf).sum().backward()
    assert rel_err(dg, (base - bias) * pf.grad) < tol
    # GELU with the DERIVATIVE as side output (saved by the forward epilogue), and the plain x * aux backward epilogue
    dpre = torch.empty(M, N, dtype=dtype, device=DEV)
    g2 = ops.linear_fwd(x, w, bias, mode=Lb.EPI_GELU_GRAD, out2=dpre)
    bf = base.clone().requires_grad_()
    F.gelu(bf).sum().backward()
    assert rel_err(g2, F.gelu(base)) < tol
    assert rel_err(dpre, bf.grad) < tol
    ma = ops.linear_fwd(x, w, None, mode=Lb.EPI_MUL_AUX, aux=pre_in)
    assert rel_err(ma, (base - bias) * pre_in.float()) < tol
    # dtanh multiply
    th = torch.tanh(rnd(M, N, seed=12)).to(dtype)
    dth = ops.linear_fwd(x, w, None, mode=Lb.EPI_MUL_DTANH, aux=th)
    assert rel_err(dth, (base - bias) * (1 - th.float() ** 2)) < tol
    # squared norm rows (+ optional store)
    rowsq = torch.zeros(M, dtype=torch.float32, device=DEV)
    tt = ops.linear_fwd(x, w, None, mode=Lb.EPI_SQNORM, rowvec=rowsq, out=torch.empty(M, N, dtype=dtype, device=DEV))
    assert rel_err(rowsq, ((base - bias) ** 2).sum(1)) < tol
    assert rel_err(tt, base - bias) < tol
    # row scale, fp32 output
    rs = rnd(M, seed=13)
    o = ops.linear_fwd(x, w, None, mode=Lb.EPI_ROWSCALE, rowvec=rs, out_dtype=torch.float32)
    assert o.dtype == torch.float32
    assert rel_err(o, (base - bias) * rs[:, None]) < tol


def test_gemm_bf16_large_persistent(ops, Lb):
    """More tiles than SMs: exercises the persistent loop, both TMEM accumulator stages and phases."""
    M, N, K = 8192, 2304, 768
    x = rnd(M, K, seed=21, dtype=torch.bfloat16)
    w = rnd(N, K, seed=22, scale=0.05, dtype=torch.bfloat16)
    y = ops.linear_fwd(x, w, None)
    ref = (x.float() @ w.float().t())
    assert rel_err(y, ref) < 1e-2


@pytest.mark.parametrize("M,N,K", [(5000, 800, 768), (4096, 2304, 768), (1111, 768, 3072), (2048, 72, 256)])
def test_gemm_pair_kernel_matches_single(ops, Lb, M, N, K):
    """CTA-pair (cta_group::2) kernel with the TMA-staged epilogue vs the single-CTA kernel with direct stores:
    ragged M / N, many tiles per cluster, every fused epilogue of the training step, same dropout mask."""
    bf = torch.bfloat16
    x = rnd(M, K, seed=31, dtype=bf)
    w = rnd(N, K, seed=32, scale=0.05, dtype=bf)
    bias = rnd(N, seed=33)
    res = rnd(M, N, seed=34, dtype=bf)
    dy = rnd(M, N, seed=35, dtype=bf)
    ref = x.float() @ w.float().t() + bias
    out = {}
    try:
        for impl in ("single", "auto"):
            ops.set_gemm_impl(impl)
            pre = torch.empty(M, N, dtype=bf, device=DEV)
            o = dict(store=ops.linear_fwd(x, w, bias),
                     gelu=ops.linear_fwd(x, w, bias, mode=Lb.EPI_GELU, out2=pre), pre=pre,
                     tanh=ops.linear_fwd(x, w, bias, mode=Lb.EPI_TANH),
                     resid=ops.linear_fwd(x, w, bias, mode=Lb.EPI_RESID, aux=res),
                     drop=ops.linear_fwd(x, w, bias, mode=Lb.EPI_RESID, aux=res, p_drop=0.1, seed=4321),
                     dgelu=ops.linear_fwd(x, w, None, mode=Lb.EPI_MUL_DGELU, aux=res),
                     dtanh=ops.linear_fwd(x, w, None, mode=Lb.EPI_MUL_DTANH, aux=torch.tanh(res.float()).to(bf)),
                     dgrad=ops.linear_dgrad(dy, w),
                     f32=ops.linear_fwd(x, w, bias, out_dtype=torch.float32))
            dw = torch.zeros(N, K, device=DEV)
            ops.linear_wgrad(dy, x, dw)
            o["wgrad"] = dw
            out[impl] = o
    finally:
        ops.set_gemm_impl("auto")
    assert rel_err(out["auto"]["store"], ref) < 1e-2
    assert rel_err(out["auto"]["gelu"], F.gelu(ref)) < 1e-2
    assert rel_err(out["auto"]["resid"], ref + res.float()) < 1e-2
    assert rel_err(out["auto"]["wgrad"], dy.float().t() @ x.float()) < 1e-2
    for k in out["auto"]:
        a, b = out["auto"][k], out["single"][k]
        assert torch.isfinite(a.float()).all(), k
        assert rel_err(a, b) < 1e-2, k
    # identical dropout masks in both kernels
    za, zb = out["auto"]["drop"].float() == res.float(), out["single"]["drop"].float() == res.float()
    assert torch.equal(za, zb)
    assert abs(float((~za).float().mean()) - 0.9) < 0.01


# ---------------------------------------------------------------------------------------- LN / embeddings
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H", [768, 1024])
def test_layernorm_fwd_bwd(ops, dtype, H):
    rows = 777
    z = rnd(rows, H, seed=1, dtype=dtype)
    gamma, beta = 1 + 0.1 * rnd(H, seed=2), 0.1 * rnd(H, seed=3)
    y, mean, rstd = ops.layernorm_fwd(z, gamma, beta, 1e-5)
    zf = z.float().requires_grad_()
    gf, bf_ = gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    ref = F.layer_norm(zf, (H,), gf, bf_, 1e-5)
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    assert rel_err(y, ref) < tol
    dy = rnd(rows, H, seed=4, dtype=dtype)
    ref.backward(dy.float())
    dg, db = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    dz, _ = ops.layernorm_bwd(dy, z, gamma, mean, rstd, dg, db)
    assert rel_err(dz, zf.grad) < tol
    assert rel_err(dg, gf.grad) < (2e-2 if dtype == torch.bfloat16 else 1e-4)
    assert rel_err(db, bf_.grad) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_layernorm_bwd_fused_dropout_and_bias_grad(ops, dtype, p_drop):
    """The fused tail: dd = dropout-masked dz (same mask as ops.dropout_apply / the RESID epilogue) and
    d_bias += column sums of dd."""
    rows, H = 1500, 768
    z = rnd(rows, H, seed=11, dtype=dtype)
    gamma, beta = 1 + 0.1 * rnd(H, seed=12), 0.1 * rnd(H, seed=13)
    _, mean, rstd = ops.layernorm_fwd(z, gamma, beta, 1e-5)
    dy = rnd(rows, H, seed=14, dtype=dtype)
    dg, db, dbias = (torch.zeros(H, device=DEV) for _ in range(3))
    dg0, db0 = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    dz_plain, same = ops.layernorm_bwd(dy, z, gamma, mean, rstd, dg0, db0)
    assert same is dz_plain
    dz, dd = ops.layernorm_bwd(dy, z, gamma, mean, rstd, dg, db, d_bias=dbias, p_drop=p_drop, seed=77)
    assert torch.equal(dz, dz_plain)
    assert rel_err(dg, dg0) < 1e-5 and rel_err(db, db0) < 1e-5     # atomics across blocks: order varies
    if p_drop > 0:
        keep = ops.dropout_apply(torch.ones(rows, H, device=DEV, dtype=dtype), p_drop, 77).float() > 0
        want = torch.where(keep, dz.float() / (1 - p_drop), torch.zeros((), device=DEV))
        assert rel_err(dd, want) < (1e-2 if dtype == torch.bfloat16 else 1e-6)
        assert torch.equal(dd.float() != 0, keep & (dz.float() != 0))
    else:
        assert dd is dz
    assert rel_err(dbias, dd.float().sum(0)) < (2e-2 if dtype == torch.bfloat16 else 1e-4)


@pytest.mark.parametrize("kind", ["roberta", "bert"])
def test_embed_ln_fwd_bwd(ops, kind):
    cfg = O.EncoderCfg.roberta_base(vocab_size=500) if kind == "roberta" else O.EncoderCfg.bert_base(vocab_size=500)
    H = cfg.hidden_size
    g = torch.Generator().manual_seed(5)
    B, Lq = 5, 37
    ids = torch.randint(0, 500, (B, Lq), generator=g)
    ids[:, 25:] = 0
    ids[1, 3] = 1
    ids[2, :4] = 1
    tts = torch.randint(0, cfg.type_vocab_size, (B, Lq), generator=g)
    p = {"bert.embeddings.word_embeddings.weight": torch.randn(500, H, generator=g) * 0.05,
         "bert.embeddings.position_embeddings.weight": torch.randn(cfg.max_position_embeddings, H, generator=g) * 0.05,
         "bert.embeddings.token_type_embeddings.weight": torch.randn(cfg.type_vocab_size, H, generator=g) * 0.05,
         "bert.embeddings.LayerNorm.weight": 1 + 0.1 * torch.randn(H, generator=g),
         "bert.embeddings.LayerNorm.bias": 0.1 * torch.randn(H, generator=g)}
    p = {k: v.requires_grad_() for k, v in p.items()}
    ref, ref_pos = O.embeddings(p, cfg, ids, tts)
    dout = torch.randn(B, Lq, H, generator=g)
    ref.backward(dout)
    d = {k: v.detach().to(DEV) for k, v in p.items()}
    kid = 0 if kind == "roberta" else 1
    out, pids, mean, rstd = ops.embed_ln_fwd(ids.to(DEV), tts.to(DEV), d["bert.embeddings.word_embeddings.weight"],
                                             d["bert.embeddings.position_embeddings.weight"],
                                             d["bert.embeddings.token_type_embeddings.weight"],
                                             d["bert.embeddings.LayerNorm.weight"], d["bert.embeddings.LayerNorm.bias"],
                                             cfg.layer_norm_eps, kid, cfg.pad_token_id, torch.float32)
    assert torch.equal(pids.cpu(), ref_pos)          # bit-exact int64
    assert rel_err(out.view(B, Lq, H), ref) < 1e-5
    grads = [torch.zeros_like(d[k]) for k in ("bert.embeddings.word_embeddings.weight",
                                              "bert.embeddings.position_embeddings.weight",
                                              "bert.embeddings.token_type_embeddings.weight",
                                              "bert.embeddings.LayerNorm.weight", "bert.embeddings.LayerNorm.bias")]
    ops.embed_ln_bwd(dout.to(DEV).view(B * Lq, H), ids.to(DEV), tts.to(DEV), pids,
                     d["bert.embeddings.word_embeddings.weight"], d["bert.embeddings.position_embeddings.weight"],
                     d["bert.embeddings.token_type_embeddings.weight"], d["bert.embeddings.LayerNorm.weight"], mean,
                     rstd, kid, cfg.pad_token_id, *grads)
    for gr, k in zip(grads, ("bert.embeddings.word_embeddings.weight", "bert.embeddings.position_embeddings.weight",
                             "bert.embeddings.token_type_embeddings.weight", "bert.embeddings.LayerNorm.weight",
                             "bert.embeddings.LayerNorm.bias")):
        assert rel_err(gr, p[k].grad) < 1e-4, k


# ---------------------------------------------------------------------------------------- attention
def _attn_ref(qkv, kp, vp, mask, B, Lq, nh, d):
    H = nh * d
    q, k, v = qkv.view(B, Lq, 3, nh, d).permute(2, 0, 3, 1, 4)
    if kp is not None:
        k = torch.cat([kp, k], 2)
        v = torch.cat([vp, v], 2)
        full = torch.cat([torch.ones(B, kp.shape[2]), mask.float()], 1)
    else:
        full = mask.float()
    s = q @ k.transpose(-1, -2) / math.sqrt(d) + O.extended_mask(full)
    pr = torch.softmax(s, -1)
    return (pr @ v).permute(0, 2, 1, 3).reshape(B * Lq, H), pr


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,Lq,P", [(2, 128, 16), (3, 40, 36), (2, 77, 0), (1, 200, 100)])
def test_attention_fwd_bwd(ops, dtype, B, Lq, P):
    nh, d = 12, 64
    g = torch.Generator().manual_seed(B * 1000 + Lq)
    qkv = (torch.randn(B * Lq, 3 * nh * d, generator=g)).to(dtype).float().requires_grad_()
    kp = vp = None
    if P:
        kp = torch.randn(B, nh, P, d, generator=g).to(dtype).float().requires_grad_()
        vp = torch.randn(B, nh, P, d, generator=g).to(dtype).float().requires_grad_()
    lens = torch.randint(5, Lq + 1, (B,), generator=g)
    mask = (torch.arange(Lq)[None] < lens[:, None]).long()
    ref, ref_pr = _attn_ref(qkv, kp, vp, mask, B, Lq, nh, d)
    dctx = torch.randn(B * Lq, nh * d, generator=g).to(dtype).float()
    ref.backward(dctx)
    to = lambda t: None if t is None else t.detach().to(dtype).to(DEV)
    ctx, lse, probs = ops.attention_fwd(to(qkv), to(kp), to(vp), mask.to(DEV), B, Lq, nh, d, want_probs=True)
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
    assert rel_err(ctx, ref) < tol
    assert rel_err(probs, ref_pr) < (1e-2 if dtype == torch.bfloat16 else 2e-5)
    dkp = torch.zeros(B, nh, P, d, device=DEV) if P else None
    dvp = torch.zeros(B, nh, P, d, device=DEV) if P else None
    dqkv = ops.attention_bwd(to(dctx), to(qkv), to(kp), to(vp), mask.to(DEV), ctx, lse, B, Lq, nh, d, dkp, dvp)
    btol = 3e-2 if dtype == torch.bfloat16 else 5e-5
    assert rel_err(dqkv, qkv.grad) < btol
    if P:
        assert rel_err(dkp, kp.grad) < btol
        assert rel_err(dvp, vp.grad) < btol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,Lq,P", [(3, 128, 16), (2, 100, 16), (3, 40, 36), (1, 200, 100), (3, 64, 16), (3, 37, 5)])
def test_attention_bwd_qkv_bias_grad(ops, dtype, B, Lq, P):
    """mtvaf_attention_bwd_ex: d_bias += column sums of dqkv on every kernel path (pipelined tcgen05 drain warps for
    L <= 128 / P <= 16, generic tcgen05 + colsum, SIMT + colsum), accumulating into what is already there."""
    nh, d = 12, 64
    H = nh * d
    qkv = rnd(B * Lq, 3 * H, seed=21, dtype=dtype)
    kp, vp = rnd(B, nh, P, d, seed=22, dtype=dtype), rnd(B, nh, P, d, seed=23, dtype=dtype)
    lens = torch.tensor([Lq, max(1, Lq // 2), max(1, Lq - 3)][:B])
    mask = (torch.arange(Lq).unsqueeze(0) < lens.unsqueeze(1)).long().to(DEV)
    ctx, lse, _ = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, nh, d, p_drop=0.1, seed=5)
    dctx = rnd(B * Lq, H, seed=24, dtype=dtype)
    base = rnd(3 * H, seed=25)
    d_bias = base.clone()
    dqkv = ops.attention_bwd(dctx, qkv, kp, vp, mask, ctx, lse, B, Lq, nh, d, p_drop=0.1, seed=5, d_bias=d_bias)
    ref = dqkv.float().sum(0)
    tol = 2e-5 if dtype == torch.float32 else 1e-2      # bf16: the fused path sums fp32 accumulators, dqkv is rounded
    assert rel_err(d_bias - base, ref) < tol
    dqkv2 = ops.attention_bwd(dctx, qkv, kp, vp, mask, ctx, lse, B, Lq, nh, d, p_drop=0.1, seed=5)
    assert torch.equal(dqkv, dqkv2)                      # the optional output does not change the gradients


@pytest.mark.parametrize("B,Lq,P,p_drop", [(2, 256, 16, 0.0), (2, 200, 36, 0.1), (3, 129, 0, 0.0), (1, 256, 100, 0.1),
                                           (2, 512, 100, 0.1), (2, 300, 16, 0.0), (3, 500, 64, 0.1), (2, 512, 0, 0.0)])
def test_attention_long_matches_simt(ops, B, Lq, P, p_drop):
    """128 < L <= 512: the tcgen05 long-text kernels (forward: resident keys, or two key windows + merge when P + L does
    not fit; backward: attention_tc_bwd_long.cu, query-tile groups) against the fp32-accumulating SIMT kernels on the
    same bf16 inputs, dropout masks included (same counter-based hash)."""
    nh, d = 12, 64
    H = nh * d
    qkv = rnd(B * Lq, 3 * H, seed=41, dtype=torch.bfloat16)
    kp = rnd(B, nh, P, d, seed=42, dtype=torch.bfloat16) if P else None
    vp = rnd(B, nh, P, d, seed=43, dtype=torch.bfloat16) if P else None
    lens = torch.tensor([Lq, max(1, Lq // 2), max(1, Lq - 5)][:B])
    mask = (torch.arange(Lq).unsqueeze(0) < lens.unsqueeze(1)).long().to(DEV)
    dctx = rnd(B * Lq, H, seed=44, dtype=torch.bfloat16)
    res = {}
    try:
        for impl in ("simt", "auto"):
            ops.set_attention_impl(impl)
            ctx, lse, _ = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, nh, d, p_drop=p_drop, seed=9)
            dkp = torch.zeros(B, nh, P, d, device=DEV) if P else None
            dvp = torch.zeros(B, nh, P, d, device=DEV) if P else None
            db = torch.zeros(3 * H, device=DEV)
            dqkv = ops.attention_bwd(dctx, qkv, kp, vp, mask, ctx, lse, B, Lq, nh, d, dkp=dkp, dvp=dvp, p_drop=p_drop,
                                     seed=9, d_bias=db)
            res[impl] = (dqkv.float(), dkp, dvp, db, ctx.float(), lse)
    finally:
        ops.set_attention_impl("auto")
    a, b = res["auto"], res["simt"]
    assert rel_err(a[4], b[4]) < 2e-2                  # forward context
    assert rel_err(a[5], b[5]) < 1e-3                  # log-sum-exp
    assert rel_err(a[0], b[0]) < 3e-2
    if P:
        assert rel_err(a[1], b[1]) < 3e-2 and rel_err(a[2], b[2]) < 3e-2
    assert rel_err(a[3], b[3]) < 3e-2


def test_attention_dropout_consistency(ops):
    """fwd/bwd regenerate the same mask: finite-difference-free check via linearity in V."""
    B, Lq, P, nh, d = 2, 64, 16, 12, 64
    qkv = rnd(B * Lq, 3 * nh * d, seed=1)
    kp, vp = rnd(B, nh, P, d, seed=2), rnd(B, nh, P, d, seed=3)
    mask = torch.ones(B, Lq, dtype=torch.long, device=DEV)
    ctx, lse, _ = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, nh, d, p_drop=0.1, seed=77)
    ctx2, _, _ = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, nh, d, p_drop=0.1, seed=77)
    assert torch.equal(ctx, ctx2)
    # ctx is linear in V for fixed probabilities+mask: <dctx, ctx(V)> gradient wrt vp must equal dvp
    dctx = rnd(B * Lq, nh * d, seed=4)
    dkp, dvp = torch.zeros_like(kp), torch.zeros_like(vp)
    ops.attention_bwd(dctx, qkv, kp, vp, mask, ctx, lse, B, Lq, nh, d, dkp, dvp, p_drop=0.1, seed=77)
    e = torch.zeros_like(vp)
    e[1, 3, 5, 7] = 1.0
    ctx_e, _, _ = ops.attention_fwd(qkv, kp, vp + e, mask, B, Lq, nh, d, p_drop=0.1, seed=77)
    fd = ((ctx_e - ctx) * dctx).sum()
    assert abs(float(fd) - float(dvp[1, 3, 5, 7])) < 1e-3 * (1 + abs(float(fd)))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_dropout_keep_rate(ops, dtype):
    """With V == 1 every context element is sum_k keep_k p_k / (1-p): its mean over many rows is 1 and the
    fraction of dropped probability mass is p (uniform probabilities: Q = K = 0)."""
    B, Lq, P, nh, d, p = 8, 128, 16, 12, 64, 0.1
    H = nh * d
    qkv = torch.zeros(B * Lq, 3 * H, device=DEV, dtype=dtype)
    qkv[:, 2 * H:] = 1.0
    kp = torch.zeros(B, nh, P, d, device=DEV, dtype=dtype)
    vp = torch.ones(B, nh, P, d, device=DEV, dtype=dtype)
    mask = torch.ones(B, Lq, dtype=torch.long, device=DEV)
    ctx, _, _ = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, nh, d, p_drop=p, seed=123)
    kept = ctx.float().view(B * Lq, nh, d)[:, :, 0] * (1 - p)          # fraction of the 144 keys kept per row
    assert abs(float(kept.mean()) - (1 - p)) < (5e-3 if dtype == torch.bfloat16 else 2e-3)   # bf16(1/144) is +0.2%
    # per-row keep counts are binomial(144, 0.9): variance of the fraction = p(1-p)/144
    var = float(kept.var())
    assert 0.5 * p * (1 - p) / (P + Lq) < var < 1.6 * p * (1 - p) / (P + Lq)
    ctx2, _, _ = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, nh, d, p_drop=p, seed=124)
    assert not torch.equal(ctx, ctx2)


@pytest.mark.parametrize("B,Lq,P,p_drop", [(2, 128, 16, 0.1), (3, 40, 36, 0.1), (2, 100, 5, 0.0), (4, 128, 64, 0.1),
                                            (2, 128, 0, 0.1), (1, 17, 4, 0.0), (2, 64, 5, 0.1), (32, 128, 16, 0.1),
                                            (13, 96, 36, 0.0), (14, 50, 16, 0.1), (30, 128, 9, 0.0)])
def test_attention_tc_matches_simt(ops, B, Lq, P, p_drop):
    """The tcgen05 forward/backward kernels and the SIMT kernels share one dropout hash: on bf16 inputs they
    must agree to bf16 rounding, with and without probability dropout, including ragged key masks."""
    nh, d = 12, 64
    bf = torch.bfloat16
    qkv = rnd(B * Lq, 3 * nh * d, seed=11, dtype=bf)
    kp = rnd(B, nh, P, d, seed=12, dtype=bf) if P else None
    vp = rnd(B, nh, P, d, seed=13, dtype=bf) if P else None
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(3, Lq + 1, (B,), generator=g)
    mask = (torch.arange(Lq)[None] < lens[:, None]).long().to(DEV)
    dctx = rnd(B * Lq, nh * d, seed=14, dtype=bf)
    res = {}
    try:
        for impl in ("simt", "tc_generic", "auto"):
            ops.set_attention_impl(impl)
            ctx, lse, _ = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, nh, d, p_drop=p_drop, seed=99)
            dkp = torch.zeros(B, nh, P, d, device=DEV) if P else None
            dvp = torch.zeros(B, nh, P, d, device=DEV) if P else None
            dqkv = ops.attention_bwd(dctx, qkv, kp, vp, mask, ctx, lse, B, Lq, nh, d, dkp, dvp, p_drop=p_drop, seed=99)
            res[impl] = (ctx, lse, dqkv, dkp, dvp)
    finally:
        ops.set_attention_impl("auto")
    names = ("ctx", "lse", "dqkv", "dkp", "dvp")
    for impl in ("tc_generic", "auto"):          # auto = software-pipelined backward when L <= 128 and P <= 16
        for nm, a, b in zip(names, res[impl], res["simt"]):
            if a is None:
                continue
            assert torch.isfinite(a.float()).all(), (impl, nm)
            assert rel_err(a, b) < 2e-2, (impl, nm)


@pytest.mark.parametrize("B,Lq,P,nh,p_drop", [(1, 64, 16, 3, 0.0), (3, 64, 16, 3, 0.1), (5, 33, 0, 1, 0.1), (2, 64, 32, 12, 0.1),
                                               (7, 9, 3, 5, 0.0)])
def test_attention_two_items_per_tile(ops, B, Lq, P, nh, p_drop):
    """L <= 64 runs two (batch, head) items per 128-row tile (attn_fwd_tc_pair_kernel; attn_bwd_pair_kernel for P <= 16):
    odd item counts (the last pair is half empty), pairs that straddle two batch rows (odd head count), ragged masks --
    forward and backward against the SIMT kernels (same dropout hash) and, without dropout, the fp32 reference."""
    d = 64
    bf = torch.bfloat16
    qkv = rnd(B * Lq, 3 * nh * d, seed=21, dtype=bf)
    kp = rnd(B, nh, P, d, seed=22, dtype=bf) if P else None
    vp = rnd(B, nh, P, d, seed=23, dtype=bf) if P else None
    g = torch.Generator().manual_seed(6)
    lens = torch.randint(1, Lq + 1, (B,), generator=g)
    mask = (torch.arange(Lq)[None] < lens[:, None]).long().to(DEV)
    dctx = rnd(B * Lq, nh * d, seed=24, dtype=bf)
    res = {}
    try:
        for impl in ("simt", "auto"):
            ops.set_attention_impl(impl)
            ctx, lse, _ = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, nh, d, p_drop=p_drop, seed=7)
            dkp = torch.zeros(B, nh, P, d, device=DEV) if P else None
            dvp = torch.zeros(B, nh, P, d, device=DEV) if P else None
            dqkv = ops.attention_bwd(dctx, qkv, kp, vp, mask, ctx, lse, B, Lq, nh, d, dkp, dvp, p_drop=p_drop, seed=7)
            res[impl] = (ctx, lse, dqkv, dkp, dvp)
    finally:
        ops.set_attention_impl("auto")
    for nm, a, b in zip(("ctx", "lse", "dqkv", "dkp", "dvp"), res["auto"], res["simt"]):
        if a is None:
            continue
        assert torch.isfinite(a.float()).all(), nm
        assert rel_err(a, b) < 2e-2, nm
    if p_drop == 0.0:
        ref, _ = _attn_ref(qkv.float().cpu(), None if kp is None else kp.float().cpu(), None if vp is None else vp.float().cpu(),
                           mask.cpu(), B, Lq, nh, d)
        assert rel_err(res["auto"][0], ref.to(DEV)) < 2e-2


# ---------------------------------------------------------------------------------------- fusion
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gate_and_mean4(ops, dtype):
    n_layers, n_img, B, hid = 12, 4, 3, 768
    W = 8 * hid
    guids = rnd(n_img, B, 4, W, seed=1, dtype=dtype)
    gf = guids.float().requires_grad_()
    # mean4 modes
    m0 = ops.mean4_fwd(guids, n_img * B, W, 0)
    assert rel_err(m0, gf.detach().view(n_img * B, 4, W).mean(1)) < (1e-2 if dtype == torch.bfloat16 else 1e-6)
    m1 = ops.mean4_fwd(guids, n_img * B, W, 1)
    ref1 = torch.stack(gf.detach().view(n_img * B, 4, W).split(2 * hid, -1)).sum(0).view(n_img * B, -1) / 4
    assert rel_err(m1, ref1) < (1e-2 if dtype == torch.bfloat16 else 1e-6)
    logits = rnd(n_img * B, n_layers * 4, seed=2).requires_grad_()
    kv, gates = ops.gate_fwd(guids, logits.detach(), n_layers, n_img, B, hid)
    # reference (bert_model.py:566-587)
    res = []
    for l in range(n_layers):
        kvs = []
        for j in range(n_img):
            sp = gf[j].split(2 * hid, -1)
            gte = torch.softmax(F.leaky_relu(logits[j * B:(j + 1) * B, l * 4:(l + 1) * 4]), -1)
            kvs.append(sum(gte[:, i].view(-1, 1, 1) * sp[i] for i in range(4)))
        c = torch.cat(kvs, 1)
        k, v = c.split(hid, -1)
        res.append(torch.stack([k.reshape(B, -1), v.reshape(B, -1)]))
    ref = torch.stack(res)                  # [L,2,B,P*hid]
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    assert rel_err(kv, ref) < tol
    dkv = rnd(*ref.shape, seed=3)
    ref.backward(dkv)
    d_guids = torch.full((n_img, B, 4, W), 7.0, device=DEV)       # gate_bwd WRITES (no pre-zeroing needed)
    d_logits = ops.gate_bwd(dkv, guids, logits.detach(), gates, n_layers, n_img, B, hid, d_guids)
    assert rel_err(d_guids, gf.grad) < tol
    assert rel_err(d_logits, logits.grad) < (2e-2 if dtype == torch.bfloat16 else 1e-4)
    # mean4 backward
    dx = torch.zeros(n_img * B, 4, W, device=DEV)
    dy = rnd(n_img * B, W, seed=4)
    ops.mean4_bwd_add(dy, dx, n_img * B, W, 0)
    assert rel_err(dx, dy[:, None, :].expand(-1, 4, -1) / 4) < 1e-6
    dx.zero_()
    ops.mean4_bwd_add(dy, dx, n_img * B, W, 1)
    x = torch.zeros(n_img * B, 4, W, device=DEV, requires_grad=True)
    (torch.stack(x.split(2 * hid, -1)).sum(0).view(n_img * B, -1) / 4 * dy).sum().backward()
    assert rel_err(dx, x.grad) < 1e-6


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_prompt_grad_combine(ops, out_dtype):
    """One-pass assembly of d(prompt): gate path + backward of both 4-way means (+ img_dropout mask)."""
    rows, hid = 6, 768
    W = 8 * hid
    d_guids = rnd(rows * 4, W, seed=1)
    d_gs = rnd(rows, W, seed=2)
    d_gm = rnd(rows, W, seed=3, dtype=out_dtype)
    ref = d_guids.clone().view(rows, 4, W)
    dx = torch.zeros(rows, 4, W, device=DEV)
    ops.mean4_bwd_add(d_gs, dx, rows, W, 1)
    p = 0.2
    gm_d = ops.dropout_apply(d_gm, p, 901)
    ops.mean4_bwd_add(gm_d.float(), dx, rows, W, 0)
    ref = ref + dx
    out = ops.prompt_grad_combine(d_guids, d_gs, d_gm, p, 901, rows, W, out_dtype)
    assert out.dtype == out_dtype
    assert rel_err(out, ref.view(rows * 4, W)) < (1e-2 if out_dtype == torch.bfloat16 else 1e-6)
    out2 = ops.prompt_grad_combine(d_guids, None, None, 0.0, 0, rows, W, torch.float32)
    assert torch.equal(out2, d_guids)


@pytest.mark.parametrize("M,N,K", [(1024, 48, 6144), (4099, 11, 768), (7, 3, 64), (300, 16, 1024)])
def test_skinny_linear(ops, M, N, K):
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3)
    y = ops.skinny_linear(x, w, b)
    ref = x.double() @ w.double().t() + b.double()
    assert rel_err(y, ref) < 1e-5
    assert torch.equal(y, ops.skinny_linear(x, w, b))          # bitwise reproducible
    assert rel_err(ops.skinny_linear(x, w), x.double() @ w.double().t()) < 1e-5


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(4099, 11, 768), (50, 12, 1024), (300, 3, 64)])
def test_skinny_linear_backward(ops, out_dtype, M, N, K):
    dy, w, x = rnd(M, N, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(M, K, seed=3)
    dx = ops.skinny_linear_dgrad(dy, w, out_dtype)
    tol = 1e-2 if out_dtype == torch.bfloat16 else 1e-5
    assert dx.dtype == out_dtype and rel_err(dx, dy.double() @ w.double()) < tol
    # with the head's input dropout: same mask as dropout_apply on the [M, K] activation
    dxd = ops.skinny_linear_dgrad(dy, w, out_dtype, 0.1, 4242)
    keep = ops.dropout_apply(torch.ones(M, K, device=DEV), 0.1, 4242) > 0
    want = torch.where(keep, (dy.double() @ w.double()) / 0.9, torch.zeros((), device=DEV, dtype=torch.float64))
    assert rel_err(dxd, want) < tol
    dw = torch.full((N, K), 0.25, device=DEV)
    ops.skinny_linear_wgrad(dy, x, dw)
    assert rel_err(dw, dy.double().t() @ x.double() + 0.25) < 1e-5


def test_softmax_kl(ops):
    B, n, heads = 5, 2089, 4
    ld = 2096
    logits_full = torch.zeros(heads * B, ld, device=DEV)
    logits_full[:, :n] = rnd(heads * B, n, seed=1, scale=2.0)
    target = torch.softmax(rnd(B, n, seed=2), -1)
    lg = logits_full[:, :n].clone().requires_grad_()
    ref = torch.stack([F.kl_div(torch.softmax(lg[h * B:(h + 1) * B], -1).log(), target, reduction="batchmean")
                       for h in range(heads)])
    ref.sum().backward()
    loss, dl = ops.softmax_kl(logits_full, n, target, B, True)
    assert rel_err(loss, ref) < 1e-5
    assert rel_err(dl[:, :n], lg.grad) < 1e-4


# ---------------------------------------------------------------------------------------- probe
def test_probe_labels_bit_exact(ops, golden_dir):
    import os
    g = torch.load(os.path.join(golden_dir, "probe_kat.pt"), weights_only=False)
    for key in ("label_kat", "label_big"):
        out = ops.probe_labels(g[key + "_in"].to(DEV).contiguous())
        assert torch.equal(out.cpu(), g[key + "_out"])
    gen = torch.Generator().manual_seed(3)
    for Lq in (1, 2, 3, 31, 128, 200, 500):
        v = torch.rand(4, Lq, generator=gen) * 60
        v[0] = torch.round(v[0])                      # many exact ties
        assert torch.equal(ops.probe_labels(v.to(DEV)).cpu(), O.construct_label(v))


def test_mse_and_pairwise(ops):
    a, b = rnd(16, 128, seed=1, scale=30), rnd(16, 128, seed=2, scale=30)
    loss, da = ops.mse(a, b, True)
    af = a.clone().requires_grad_()
    ref = F.mse_loss(af, b)
    ref.backward()
    assert rel_err(loss, ref.view(1)) < 1e-5
    assert rel_err(da, af.grad) < 1e-5
    T = rnd(2 * 50, 384, seed=3)
    T[7] = T[9]                                        # identical rows -> exactly 0
    dist = ops.pairwise_sqdist(T, 2, 50, 384)
    t = T.view(2, 50, 384)
    ref = ((t.unsqueeze(2) - t.unsqueeze(1)) ** 2).sum(-1)
    assert rel_err(dist, ref) < 1e-5
    assert float(dist[0].diagonal().abs().max()) == 0.0 and float(dist[0, 7, 9]) == 0.0
    assert torch.equal(dist, dist.transpose(1, 2))


@pytest.mark.parametrize("B,Lq,R", [(3, 64, 384), (2, 128, 384), (2, 200, 512), (2, 512, 384), (1, 130, 64)])
def test_pairwise_sqdist_tcgen05_gram_form(ops, B, Lq, R):
    """TwoWordPSDProbe (probes/probe.py:25-46) on the tensor cores: Gram form with the bf16 hi/lo split against the
    reference's explicit differences in float64 -- <= 1e-4 (max-norm), exactly symmetric, diagonal exactly 0,
    duplicated and near-duplicated rows (cancellation) included; against the SIMT explicit-difference kernel too."""
    T = rnd(B * Lq, R, seed=300 + Lq, scale=0.7)
    T[7] = T[9]                                         # identical rows -> exactly 0
    T[Lq - 1] = T[3] * (1 + 1e-4)                       # near-duplicates: d ~ 1e-8 |t|^2, below the Gram form's digits
    T[11] = T[12] + 1e-3 * rnd(R, seed=5)
    t = T.view(B, Lq, R).double().cpu()
    ref = torch.stack([((t[b].unsqueeze(1) - t[b].unsqueeze(0)) ** 2).sum(-1) for b in range(B)])
    dist = ops.pairwise_sqdist(T, B, Lq, R)
    try:
        ops.set_pairwise_impl("simt")
        dist_simt = ops.pairwise_sqdist(T, B, Lq, R)
    finally:
        ops.set_pairwise_impl("auto")
    assert rel_err(dist_simt, ref.float()) < 1e-5
    err = rel_err(dist, ref.float())
    print("pairwise Gram form B=%d L=%d R=%d: max-norm rel err %.2e" % (B, Lq, R, err))
    assert err < 1e-4
    assert torch.equal(dist, dist.transpose(1, 2))
    assert float(dist.diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    assert float(dist[0, 7, 9]) == 0.0 and float(dist.min()) >= 0.0
    # the pairs the Gram form cannot resolve keep the reference's RELATIVE accuracy (explicit fix-up)
    for (i, j) in ((Lq - 1, 3), (11, 12)):
        r = float(ref[0, i, j])
        assert abs(float(dist[0, i, j]) - r) <= 1e-3 * r, (i, j, float(dist[0, i, j]), r)
    # element-wise relative accuracy away from the cancellation regime
    big = ref > 1e-2 * ref.max()
    assert float(((dist.cpu().double() - ref).abs() / ref.clamp_min(1e-30))[big].max()) < 1e-4


def test_pack_features_wire_format(ops):
    """mtvaf_pack_features == cat(pyramid) + view(B, 4, -1) + aux permute + cast (models/bert_model.py:536-539), from
    fp32 features and from the bf16 wire format."""
    B, n_aux = 5, 3
    img = rnd(B, 3840, 2, 2, seed=11).abs()
    aux = rnd(B, n_aux, 3840, 2, 2, seed=12).abs()
    want = torch.stack([img.reshape(B, 4, -1)] + [aux[:, j].reshape(B, 4, -1) for j in range(n_aux)])   # [n_img,B,4,3840]
    for in_dt in (torch.float32, torch.bfloat16):
        for out_dt in (torch.float32, torch.bfloat16):
            got = ops.pack_features(img.to(in_dt), aux.to(in_dt), out_dt)
            assert got.shape == (1 + n_aux, B, 3840 * 4) and got.dtype == out_dt
            assert torch.equal(got.view(1 + n_aux, B, 4, 3840), want.to(in_dt).to(out_dt))
    got = ops.pack_features(img, None, torch.bfloat16)
    assert torch.equal(got.view(1, B, 4, 3840), want[:1].to(torch.bfloat16))


# ---------------------------------------------------------------------------------------- CRF
def test_crf(ops):
    g = torch.Generator().manual_seed(4)
    B, Lq, T = 9, 40, 11
    em = torch.randn(B, Lq, T, generator=g).requires_grad_()
    start = (torch.rand(T, generator=g) * 0.2 - 0.1).requires_grad_()
    end = (torch.rand(T, generator=g) * 0.2 - 0.1).requires_grad_()
    trans = (torch.rand(T, T, generator=g) * 2 - 1).requires_grad_()
    lens = torch.randint(1, Lq + 1, (B,), generator=g)
    lens[0], lens[1] = 1, Lq
    mask = (torch.arange(Lq)[None] < lens[:, None]).long()
    tags = torch.randint(0, T, (B, Lq), generator=g)
    ref = -O.crf_log_likelihood(em, tags, mask, start, end, trans).mean()
    ref.backward()
    c = lambda t: t.detach().to(DEV).contiguous()
    nll, d_em, d_s, d_e, d_t = ops.crf_nll(c(em), c(tags), c(mask), c(start), c(end), c(trans), True, 1.0 / B)
    assert rel_err(nll / B, ref.view(1)) < 1e-5
    assert rel_err(d_em, em.grad) < 1e-4
    assert rel_err(d_s, start.grad) < 1e-4
    assert rel_err(d_e, end.grad) < 1e-4
    assert rel_err(d_t, trans.grad) < 1e-4
    best, ln = ops.crf_decode(c(em), c(mask), c(start), c(end), c(trans))
    dec = O.crf_decode(em, mask, start, end, trans)
    assert ln.cpu().tolist() == lens.tolist()
    for b in range(B):
        assert best[b, :lens[b]].cpu().tolist() == dec[b]
        assert (best[b, lens[b]:] == -1).all()


def test_combine_loss_and_adamw(ops):
    nll = torch.tensor([33.0], device=DEV)
    pl = torch.tensor([1500.0], device=DEV)
    img = torch.tensor([0.3, 0.2, 0.1, 0.4], device=DEV)
    out, flag = ops.combine_loss(nll, 4, pl, 0.5, 30, img, 0.1)
    ref = 33.0 / 4 + 1500.0 * 0.5 * 2 ** -30 + 0.1 * 1.0
    assert abs(float(out) - ref) < 1e-5 and int(flag) == 1
    out, flag = ops.combine_loss(nll, 4, torch.tensor([0.05], device=DEV), 0.5, 30, None, 0.1)
    assert abs(float(out) - 33.0 / 4) < 1e-6 and int(flag) == 0
    p = rnd(1000, seed=1).requires_grad_()
    p2 = p.detach().clone()
    opt = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m, v = torch.zeros_like(p2), torch.zeros_like(p2)
    for step in range(1, 4):
        gr = rnd(1000, seed=10 + step)
        p.grad = gr.clone()
        opt.step()
        ops.adamw_step(p2, gr, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, step)
    assert rel_err(p2, p.detach()) < 1e-6


def test_adamw_vector_path_zero_grad_and_bf16_shadow(ops):
    """16-byte vector AdamW with the fused gradient clear and bf16 weight shadow, odd length (scalar tail)."""
    n = 4 * 1000 + 3
    p = rnd(n, seed=1).requires_grad_()
    p2 = p.detach().clone()
    opt = torch.optim.AdamW([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m, v = torch.zeros_like(p2), torch.zeros_like(p2)
    shadow = torch.zeros(n, dtype=torch.bfloat16, device=DEV)
    for step in range(1, 4):
        gr = rnd(n, seed=10 + step)
        p.grad = gr.clone()
        opt.step()
        ops.adamw_step(p2, gr, m, v, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, bf16_copy=shadow, zero_grad=True)
        assert float(gr.abs().max()) == 0.0
    assert rel_err(p2, p.detach()) < 1e-6
    assert torch.equal(shadow, p2.to(torch.bfloat16))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,ld", [(1000, 768, 768), (4099, 3072, 3072), (333, 2304, 2304), (700, 2089, 2096),
                                    (513, 11, 11), (64, 256, 512)])
def test_colsum(ops, dtype, M, N, ld):
    x = rnd(M, ld, seed=3, dtype=dtype)
    db = torch.full((N,), 0.5, device=DEV)
    ops.colsum(x, db, n_valid=N)
    ref = x[:, :N].float().sum(0) + 0.5
    assert rel_err(db, ref) < 1e-4


def test_span_heads_kernels(ops):
    """distant CE, mean CE and the ragged span pooling (fwd + bwd) against the oracle's torch restatement."""
    g = torch.Generator().manual_seed(3)
    B, Lq, M, H = 4, 24, 5, 768
    lens = torch.tensor([24, 9, 17, 3])
    mask = (torch.arange(Lq)[None] < lens[:, None]).long()
    seq = torch.randn(B, Lq, H, generator=g).requires_grad_()
    starts = torch.tensor([[1, 5, 0, 0, 20], [2, 0, 0, 0, 0], [3, 10, 16, 0, 0], [1, 0, 0, 0, 0]])
    ends = torch.tensor([[3, 5, 0, 0, 23], [8, 0, 0, 0, 0], [4, 12, 16, 0, 0], [2, 0, 0, 0, 0]])
    w = (torch.randn(1, H, generator=g) * 0.05).requires_grad_()
    b = torch.tensor([0.1], requires_grad=True)
    emb, smask = O.span_representation(starts, ends, seq, mask)
    score = F.linear(emb, w, b).squeeze(-1)
    ref = O.self_att_representation(emb, score, smask)
    dpool = torch.randn(ref.shape, generator=g)
    ref.backward(dpool)
    ws = ops.span_offsets(mask.to(DEV))
    assert ws[:B].tolist() == lens.tolist() and int(ws[2 * B]) == int(lens.sum())
    sd = seq.detach().reshape(B * Lq, H).to(DEV).contiguous()
    pooled = ops.span_pool_fwd(sd, ws, starts.to(DEV), ends.to(DEV), w.detach().view(-1).to(DEV), b.detach().to(DEV),
                               B, Lq)
    assert rel_err(pooled, ref) < 1e-5
    d_seq = torch.zeros(B * Lq, H, device=DEV)
    d_w, d_b = torch.zeros(H, device=DEV), torch.zeros(1, device=DEV)
    ops.span_pool_bwd(dpool.to(DEV), sd, ws, starts.to(DEV), ends.to(DEV), w.detach().view(-1).to(DEV),
                      b.detach().to(DEV), B, Lq, d_seq, d_w, d_b)
    assert rel_err(d_seq, seq.grad.reshape(B * Lq, H)) < 1e-4
    assert rel_err(d_w, w.grad.view(-1)) < 1e-4
    assert abs(float(d_b) - float(b.grad)) < 1e-4 * (1 + abs(float(b.grad)))
    # distant cross-entropy on the two columns of a [T, 2] matrix
    ae = torch.randn(B * Lq, 2, generator=g).requires_grad_()
    pos = torch.zeros(B, Lq, dtype=torch.long)
    pos[0, 3] = pos[0, 7] = pos[1, 2] = pos[2, 5] = pos[3, 1] = 1
    l0 = O.distant_cross_entropy(ae.view(B, Lq, 2)[..., 0], pos)
    l1 = O.distant_cross_entropy(ae.view(B, Lq, 2)[..., 1], pos)
    ((l0 + l1) / 2).backward()
    loss = torch.zeros(1, device=DEV)
    d_ae = torch.empty(B * Lq, 2, device=DEV)
    ops.distant_ce(ae.detach().to(DEV), 0, pos.to(DEV), 0.5, loss, d_ae)
    ops.distant_ce(ae.detach().to(DEV), 1, pos.to(DEV), 0.5, loss, d_ae)
    assert abs(float(loss) - float((l0 + l1) / 2)) < 1e-5
    assert rel_err(d_ae, ae.grad) < 1e-5
    # mean cross-entropy
    lg = torch.randn(37, 4, generator=g).requires_grad_()
    y = torch.randint(0, 4, (37,), generator=g)
    F.cross_entropy(lg, y).backward()
    loss = torch.zeros(1, device=DEV)
    d = ops.ce_mean(lg.detach().to(DEV), y.to(DEV), 1.0, loss, True)
    assert abs(float(loss) - float(F.cross_entropy(lg, y))) < 1e-5
    assert rel_err(d, lg.grad) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_row_sqnorm(ops, dtype):
    x = rnd(1000, 384, seed=77, dtype=dtype)
    got = ops.row_sqnorm(x)
    ref = (x.float().double() ** 2).sum(-1).float()
    assert rel_err(got, ref) < 1e-6
    xs = rnd(37, 512, seed=78, dtype=dtype)[:, :128]          # strided rows
    assert rel_err(ops.row_sqnorm(xs), (xs.float().double() ** 2).sum(-1).float()) < 1e-6
