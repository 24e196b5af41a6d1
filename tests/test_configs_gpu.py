"""GPU parity at the shapes of BASELINE.json configs[2..4] (the bench line is configs[1]):

  configs[2]  roberta-large backbone (24 layers, H=1024, 16 heads, I=4096), long auxiliary text L=256, P=36 regions
  configs[3]  roberta-base inference, bs=512, L=128, P=36, no_grad  (full size: checked through batch-independence
              and a sub-batch compared with the oracle)
  configs[4]  fusion-layer + psdProbe microbenchmark grid: L in 64..512 x regions in 10..100 -- attention forward /
              backward against the explicit softmax formula (models/modeling_roberta.py:191-284), OneWord / TwoWord
              probe (layer-4 / layer-7 shaped 768x384 projections)

Tolerances: fp32 <= 1e-4 relative, bf16 <= 2e-2 relative (north_star), index work bit-exact."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mtvaf_oracle as O
from oracle.make_golden import hf_config
from mtvaf_b200 import synthetic as S

DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _encoder(cfg, params, dtype):
    from mtvaf_b200.modules import RobertaModel
    m = RobertaModel.from_config(hf_config(cfg), compute_dtype=dtype)
    sd = {k[len("bert."):]: v for k, v in params.items() if k.startswith("bert.")}
    m.load_state_dict(sd, strict=False)
    return m.to(DEV).eval()


# ------------------------------------------------------------------------------------------------ configs[2]
@pytest.mark.parametrize("dtype,tol,gtol", [("fp32", 1e-4, 2e-3), ("bf16", 2e-2, 8e-2)])
def test_roberta_large_long_text_encoder_matches_oracle(dtype, tol, gtol):
    cfg = O.EncoderCfg.roberta_large(vocab_size=3000)
    B, Lq, P = 2, 256, 36
    params = S.init_params(cfg, seed=31, ln_jitter=0.05, with_fusion=False)
    batch = S.make_batch(B, Lq, vocab=cfg.vocab_size, shape="longaux", seed=32, with_images=False)
    pkv = S.make_prefix(B, cfg.num_hidden_layers, cfg.num_attention_heads, P, 64, seed=33)
    mask = torch.cat([torch.ones(B, P), batch["attention_mask"].float()], dim=1)

    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}
    okv = [(k.clone().requires_grad_(), v.clone().requires_grad_()) for k, v in pkv]
    o = O.encoder_forward(p, cfg, batch["input_ids"], mask, batch["token_type_ids"], past_key_values=okv)
    gen = torch.Generator().manual_seed(5)
    w = torch.randn(o["last_hidden_state"].shape, generator=gen)
    (o["last_hidden_state"] * w).sum().backward()

    m = _encoder(cfg, params, dtype)
    dkv = [(k.to(DEV).requires_grad_(), v.to(DEV).requires_grad_()) for k, v in pkv]
    enc = m(input_ids=batch["input_ids"].to(DEV), attention_mask=mask.to(DEV),
            token_type_ids=batch["token_type_ids"].to(DEV), past_key_values=dkv, output_attentions=True,
            output_hidden_states=True, return_dict=True)
    hs = enc["hidden_states"]
    assert len(hs) == cfg.num_hidden_layers + 1
    assert rel(hs[0], o["hidden_states"][0]) < tol
    assert rel(hs[12], o["hidden_states"][12]) < tol
    assert rel(hs[-1], o["last_hidden_state"]) < tol
    (hs[-1].float() * w.to(DEV)).sum().backward()
    for name in ("encoder.layer.23.output.dense.weight", "encoder.layer.11.attention.self.query.weight",
                 "encoder.layer.0.intermediate.dense.weight", "embeddings.position_embeddings.weight"):
        got = dict(m.named_parameters())[name].grad.cpu()
        ref = p["bert." + name].grad
        assert float((got - ref).norm() / ref.norm()) < gtol, name
    for i in (0, 23):
        for j in (0, 1):
            ref = okv[i][j].grad
            assert float((dkv[i][j].grad.cpu() - ref).norm() / ref.norm()) < gtol, (i, j)


# ------------------------------------------------------------------------------------------------ configs[3]
def test_inference_bs512_batch_independent_and_matches_oracle():
    """roberta-base eval, bs=512, L=128, P=36, bf16, no_grad: every sample's output must not depend on its
    batch-mates (size-independent property at the full size), and a sub-batch must match the oracle."""
    cfg = O.EncoderCfg.roberta_base(vocab_size=4000)
    B, Lq, P, sub = 512, 128, 36, 4
    params = S.init_params(cfg, seed=41, ln_jitter=0.05, with_fusion=False)
    batch = S.make_batch(B, Lq, vocab=cfg.vocab_size, shape="twitter2017", seed=42, with_images=False)
    pkv = S.make_prefix(B, cfg.num_hidden_layers, cfg.num_attention_heads, P, 64, seed=43)
    mask = torch.cat([torch.ones(B, P), batch["attention_mask"].float()], dim=1)
    m = _encoder(cfg, params, "bf16")

    def run(sl):
        with torch.no_grad():
            enc = m(input_ids=batch["input_ids"][sl].to(DEV), attention_mask=mask[sl].to(DEV),
                    token_type_ids=batch["token_type_ids"][sl].to(DEV),
                    past_key_values=[(k[sl].to(DEV), v[sl].to(DEV)) for k, v in pkv],
                    output_attentions=True, output_hidden_states=True, return_dict=True)
        return enc
    full = run(slice(0, B))
    assert full["last_hidden_state"].shape == (B, Lq, cfg.hidden_size)
    assert not full["last_hidden_state"].requires_grad
    assert torch.isfinite(full["last_hidden_state"].float()).all()
    for sl in (slice(0, sub), slice(B - sub, B), slice(255, 255 + sub)):
        part = run(sl)
        # same kernels, same per-row arithmetic: bit-identical regardless of the batch around the sample
        assert torch.equal(part["last_hidden_state"], full["last_hidden_state"][sl])
        assert torch.equal(part["hidden_states"][7], full["hidden_states"][7][sl])
    p = {k: v for k, v in params.items()}
    with torch.no_grad():
        o = O.encoder_forward(p, cfg, batch["input_ids"][:sub], mask[:sub], batch["token_type_ids"][:sub],
                              past_key_values=[(k[:sub], v[:sub]) for k, v in pkv])
    assert rel(full["last_hidden_state"][:sub].float(), o["last_hidden_state"]) < 2e-2
    assert rel(full["hidden_states"][4][:sub].float(), o["hidden_states"][4]) < 2e-2


# ------------------------------------------------------------------------------------------------ configs[4]
def _attn_reference(qkv, kp, vp, key_mask, B, Lq, nh, d):
    """softmax(Q [K_p;K]^T / sqrt(d) + (1-mask) * -10000) [V_p;V]  (models/modeling_roberta.py:191-284), fp32."""
    H = nh * d
    q, k, v = (qkv[:, i * H:(i + 1) * H].view(B, Lq, nh, d).permute(0, 2, 1, 3) for i in range(3))
    P = 0 if kp is None else kp.shape[2]
    if P:
        k = torch.cat([kp, k], dim=2)
        v = torch.cat([vp, v], dim=2)
    add = torch.cat([torch.zeros(B, P, device=qkv.device), (1.0 - key_mask.float()) * -10000.0], dim=1)
    s = q @ k.transpose(-1, -2) / math.sqrt(d) + add[:, None, None, :]
    ctx = torch.softmax(s, dim=-1) @ v
    return ctx.permute(0, 2, 1, 3).reshape(B * Lq, H)


SWEEP = [(2, 64, 10), (2, 64, 100), (2, 256, 36), (1, 256, 100), (1, 512, 10), (1, 512, 64), (1, 512, 100)]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("B,Lq,P", SWEEP)
def test_sweep_attention_fwd_bwd(B, Lq, P, dtype, tol):
    from mtvaf_b200 import ops
    nh, d = 12, 64
    H = nh * d
    g = torch.Generator().manual_seed(Lq * 1000 + P)
    qkv32 = torch.randn(B * Lq, 3 * H, generator=g).to(DEV)
    kp32 = torch.randn(B, nh, P, d, generator=g).to(DEV)
    vp32 = torch.randn(B, nh, P, d, generator=g).to(DEV)
    lens = torch.randint(Lq // 3, Lq + 1, (B,), generator=g)
    key_mask = (torch.arange(Lq).unsqueeze(0) < lens.unsqueeze(1)).long().to(DEV)
    w = torch.randn(B * Lq, H, generator=g).to(DEV)
    # reference on the values the kernel sees (bf16-rounded inputs in bf16 mode)
    qkv_r = qkv32.to(dtype).float().requires_grad_()
    kp_r = kp32.to(dtype).float().requires_grad_()
    vp_r = vp32.to(dtype).float().requires_grad_()
    ref = _attn_reference(qkv_r, kp_r, vp_r, key_mask, B, Lq, nh, d)
    (ref * w.to(dtype).float()).sum().backward()

    qkv, kp, vp = qkv32.to(dtype), kp32.to(dtype).contiguous(), vp32.to(dtype).contiguous()
    ctx, lse, _ = ops.attention_fwd(qkv, kp, vp, key_mask, B, Lq, nh, d)
    assert rel(ctx.float(), ref) < tol
    dkp = torch.zeros(B, nh, P, d, device=DEV)
    dvp = torch.zeros(B, nh, P, d, device=DEV)
    dqkv = ops.attention_bwd(w.to(dtype), qkv, kp, vp, key_mask, ctx, lse, B, Lq, nh, d, dkp=dkp, dvp=dvp)
    btol = 1e-4 if dtype == torch.float32 else 3e-2
    assert rel(dqkv.float(), qkv_r.grad) < btol
    assert rel(dkp, kp_r.grad) < btol
    assert rel(dvp, vp_r.grad) < btol


@pytest.mark.parametrize("layer", [4, 7])
@pytest.mark.parametrize("Lq", [64, 512])
def test_sweep_probes(layer, Lq):
    """OneWord (probes/probe.py:62-79) and TwoWord (:25-46) probes at the sweep's shortest and longest text lengths.
    The layer-4 / layer-7 matrices the reference ships do not travel to the GPU box (tests/golden/probe_kat.pt pins
    the oracle against them on CPU); here `proj` is the reference's own init U(-0.05, 0.05) (probes/probe.py:60),
    one seed per layer."""
    from mtvaf_b200 import ops
    B, H, R = 2, 768, 384
    g = torch.Generator().manual_seed(layer * 100 + Lq)
    proj = torch.rand(H, R, generator=g) * 0.1 - 0.05
    x = torch.randn(B, Lq, H, generator=g)
    ref1 = O.one_word_psd_probe(x, proj)
    ref2 = O.two_word_psd_probe(x[:, :64], proj)            # the explicit [B,L,L,r] difference tensor: keep it small
    ref2_full = torch.stack([O.two_word_psd_probe(x[b:b + 1], proj)[0] for b in range(B)]) if Lq <= 128 else None
    xd, pd = x.to(DEV), proj.to(DEV)
    T = ops.linear_fwd(xd.view(B * Lq, H), pd.t().contiguous(), None)
    norms = (T * T).sum(-1).view(B, Lq)
    assert rel(norms, ref1) < 1e-4
    T64 = ops.linear_fwd(xd[:, :64].reshape(B * 64, H), pd.t().contiguous(), None)
    D = ops.pairwise_sqdist(T64, B, 64, T64.shape[1])
    assert rel(D, ref2) < 1e-4
    assert torch.equal(D, D.transpose(1, 2)) and float(D.diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    # the full text length through the drop-in module (tcgen05 Gram form, several 128-token tiles at L=512)
    from mtvaf_b200.modules import TwoWordPSDProbe
    tw = TwoWordPSDProbe({"probe": {"maximum_rank": R}, "model": {"hidden_dim": H}}).to(DEV)
    with torch.no_grad():
        tw.proj.copy_(pd)
    Dm = tw(xd)
    assert Dm.shape == (B, Lq, Lq) and torch.equal(Dm, Dm.transpose(1, 2))
    assert float(Dm.diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    assert rel(Dm[:, :64, :64], ref2) < 1e-4
    if ref2_full is not None:
        assert rel(Dm, ref2_full) < 1e-4
    # pseudo labels from the depth norms: integer-valued, bit-exact against the restated rule
    lab = ops.probe_labels(norms)
    assert torch.equal(lab.cpu(), O.construct_label(norms.cpu()))


# ------------------------------------------------------------------------------------------------ --use_align long text
@pytest.mark.parametrize("dtype,tol,gtol", [("fp32", 1e-4, 5e-3), ("bf16", 2e-2, 8e-2)])
def test_tvnet2_long_aligned_text_l500_matches_oracle(dtype, tol, gtol):
    """The reference's `--use_align` inputs: caption + OCR + face + ANP text concatenated behind the tweet, up to
    max_seq_agn = 500 tokens (MTVAF_training.py:250,342-348; modules/dataset.py:241-261), through the whole
    TVNetSAModel2 with the visual prefix (P = 16).  In bf16 this is the long-text tcgen05 path end to end: forward as two
    key windows + merge (16 + 500 keys do not fit one resident tile set), backward with two query-tile groups."""
    from types import SimpleNamespace
    from mtvaf_b200.modules import TVNetSAModel2, FeatureStub
    cfg = O.EncoderCfg.roberta_base(vocab_size=1500)
    B, Lq = 2, 500
    params = S.init_params(cfg, seed=51, ln_jitter=0.05)
    batch = S.make_batch(B, Lq, vocab=cfg.vocab_size, shape="longaux", seed=52)
    batch["attention_mask"][0, 470:] = 0                      # ragged: one sequence shorter than the other
    batch["input_ids"][0, 470:] = 0
    batch["labels"][0, 469] = 10
    batch["labels"][0, 470:] = 0
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}
    o = O.tvnet2_forward(p, cfg, batch, alpha=0.1, beta=0.5)
    o["loss"].backward()
    args = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True, beta=0.5,
                           alpha=0.1, vao=True, noauxloss=False, resnet_root=None, compute_dtype=dtype, probe_ckpt="")
    m = TVNetSAModel2(list(range(10)), None, args, config=hf_config(cfg), image_model=FeatureStub())
    m.load_state_dict(params, strict=False)
    m = m.to(DEV).eval()
    out, prob, img = m(**{k: v.to(DEV) for k, v in batch.items()})
    assert rel(out.loss, o["loss"]) < tol
    assert rel(m.last_emissions, o["emissions"]) < tol
    assert rel(prob, o["prob_loss"]) < tol
    if dtype == "fp32":
        assert out.logits == o["logits"]
    out.loss.backward()
    for name in ("bert.encoder.layer.11.attention.self.value.weight", "bert.encoder.layer.0.attention.self.key.weight",
                 "bert.encoder.layer.5.intermediate.dense.weight", "encoder_conv.2.weight", "fc.weight"):
        got = dict(m.named_parameters())[name].grad.cpu()
        ref = p[name].grad
        assert float((got - ref).norm() / ref.norm()) < gtol, name
