"""CPU: the oracle (oracle/mtvaf_oracle.py) against the golden vectors produced by the unmodified
reference (oracle/make_golden.py) -- this is what pins the oracle on the GPU box, where
/root/reference does not exist."""
import os

import pytest
import torch

from oracle import mtvaf_oracle as O
from oracle.make_golden import CASES, ocfg_for, grad_fingerprint
from mtvaf_b200 import synthetic as S


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _close(a, b, rtol=2e-5, atol=2e-5):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def _check_fp(fp, gold_fp, rtol):
    """grad fingerprints = [l2 norm, sum, first 4 elements]; the sum is cancellation-prone so it is
    compared loosely; true-zero gradients (e.g. key bias: softmax is shift invariant) are noise."""
    for k, ref in gold_fp.items():
        if ref is None:
            assert fp[k] is None or float(fp[k][0]) == 0.0
            continue
        n = float(ref[0])
        if n < 1e-4:
            assert float(fp[k][0]) < 1e-3, k
            continue
        assert abs(float(fp[k][0]) - n) <= rtol * n, (k, fp[k], ref)
        torch.testing.assert_close(fp[k][2:], ref[2:], rtol=rtol, atol=rtol * n, msg=lambda m: k + ": " + m)


def test_label_known_answers(golden_dir):
    g = _load(golden_dir, "probe_kat")
    assert torch.equal(O.construct_label(g["label_kat_in"]), g["label_kat_out"])
    assert g["label_kat_out"].tolist() == [[4, 2, 1, 3, 2, 3], [2, 3, 1, 5, 4, 3]]
    assert torch.equal(O.construct_label(g["label_big_in"]), g["label_big_out"])


def test_position_ids_bit_exact():
    ids = torch.tensor([[0, 5, 1, 7, 0, 0], [1, 1, 3, 4, 5, 0]])
    # pad id 0 != padding_idx 1: zeros keep counting
    assert O.roberta_position_ids(ids).tolist() == [[2, 3, 1, 4, 5, 6], [1, 1, 2, 3, 4, 5]]


@pytest.mark.parametrize("name", ["encoder_roberta_p36", "encoder_bert"])
def test_encoder_matches_reference_golden(golden_dir, name):
    g = _load(golden_dir, name)
    c = CASES[name]
    cfg = ocfg_for(c["kind"])
    params = S.init_params(cfg, seed=c["param_seed"], ln_jitter=0.05, with_fusion=False)
    params = {k: v.requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}
    batch = S.make_batch(c["B"], c["L"], vocab=cfg.vocab_size, shape=c["shape"], seed=c["batch_seed"],
                         with_images=False)
    P = c.get("P", 0)
    mask = batch["attention_mask"].float()
    pkv = None
    if P:
        pkv = S.make_prefix(c["B"], cfg.num_hidden_layers, cfg.num_attention_heads, P, 64, seed=c["batch_seed"] + 1000)
        pkv = [(k.requires_grad_(), v.requires_grad_()) for k, v in pkv]
        mask = torch.cat([torch.ones(c["B"], P), mask], dim=1)
    tt = batch["token_type_ids"]
    if c["kind"] == "bert":
        tt = (torch.arange(c["L"]).unsqueeze(0).expand(c["B"], -1) % 2).contiguous()
    enc = O.encoder_forward(params, cfg, batch["input_ids"], mask, tt, pkv)
    hs = enc["hidden_states"]
    _close(hs[-1], g["last"])
    _close(hs[7][:, :, :16], g["hs7_slice"])
    _close(hs[0][:, :, :16], g["emb_slice"])
    _close(O.pooler(params, hs[-1]), g["pooler"])
    gen = torch.Generator().manual_seed(99)
    w7 = torch.randn(hs[7].shape, generator=gen)
    w12 = torch.randn(hs[12].shape, generator=gen)
    obj = (hs[7] * w7).sum() + (hs[12] * w12).sum()
    _close(obj, g["obj"], rtol=1e-4, atol=1e-3)
    obj.backward()
    fp = grad_fingerprint([(k, v.grad) for k, v in params.items() if k.startswith("bert.")])
    _check_fp(fp, g["grad_fp"], 2e-4)
    if P:
        _close(pkv[0][0].grad, g["dk0"], rtol=1e-4, atol=1e-5)


def test_tvnet2_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "tvnet2_roberta")
    c = CASES["tvnet2_roberta"]
    cfg = ocfg_for(c["kind"])
    params = S.init_params(cfg, seed=c["param_seed"], ln_jitter=0.05)
    params = {k: v.requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}
    batch = S.make_batch(c["B"], c["L"], vocab=cfg.vocab_size, shape=c["shape"], seed=c["batch_seed"])
    o = O.tvnet2_forward(params, cfg, batch, alpha=0.1, beta=0.5)
    _close(o["loss"], g["loss"], rtol=1e-5, atol=1e-5)
    _close(o["prob_loss"], g["prob_loss"], rtol=1e-5, atol=1e-3)
    _close(o["img_loss"], g["img_loss"], rtol=1e-5, atol=1e-6)
    assert o["logits"] == g["logits"]
    _close(o["emissions"], g["emissions"])
    _close(o["norms"], g["norms"], rtol=1e-5, atol=1e-3)
    assert torch.equal(o["pseudo_labels"], g["pseudo_labels"])
    _close(o["prefix"][0][0], g["prefix_k0"])
    _close(o["prefix"][11][1][:, :, :, :8], g["prefix_v11_slice"])
    o["loss"].backward()
    fp = grad_fingerprint([(k, v.grad) for k, v in params.items()])
    _check_fp(fp, g["grad_fp"], 5e-4)


def test_tvnet2_bert_backbone_matches_reference_golden(golden_dir):
    """The BERT branch of TVNetSAModel2 (models/bert_model.py:425-429 picks the backbone by name; token-type
    embeddings, eps 1e-12, absolute positions 0..L-1): oracle vs the unmodified reference."""
    g = _load(golden_dir, "tvnet2_bert")
    c = CASES["tvnet2_bert"]
    cfg = ocfg_for(c["kind"])
    params = S.init_params(cfg, seed=c["param_seed"], ln_jitter=0.05)
    params = {k: v.requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}
    batch = S.make_batch(c["B"], c["L"], vocab=cfg.vocab_size, shape=c["shape"], seed=c["batch_seed"])
    o = O.tvnet2_forward(params, cfg, batch, alpha=0.1, beta=0.5)
    _close(o["loss"], g["loss"], rtol=1e-5, atol=1e-5)
    _close(o["prob_loss"], g["prob_loss"], rtol=1e-5, atol=1e-3)
    _close(o["img_loss"], g["img_loss"], rtol=1e-5, atol=1e-6)
    assert o["logits"] == g["logits"]
    _close(o["emissions"], g["emissions"])
    assert torch.equal(o["pseudo_labels"], g["pseudo_labels"])
    o["loss"].backward()
    fp = grad_fingerprint([(k, v.grad) for k, v in params.items()])
    _check_fp(fp, g["grad_fp"], 5e-4)


@pytest.mark.parametrize("variant", ["noauxloss", "no_vao", "no_probe", "no_prefix"])
def test_tvnet2_flag_variants_match_reference_golden(golden_dir, variant):
    """The flag branches of TVNetSAModel2.forward (noauxloss :489, vao :549-563, use_probe :527-532, use_prefix :486-492)
    through the unmodified reference (oracle/make_variant_golden.py) vs the oracle."""
    gold = _load(golden_dir, "tvnet2_variants")
    c, g = gold["case"], gold["variants"][variant]
    cfg = O.EncoderCfg.roberta_base(vocab_size=c["vocab"])
    params = S.init_params(cfg, seed=c["param_seed"], ln_jitter=0.05)
    params = {k: v.requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}
    batch = S.make_batch(c["B"], c["L"], vocab=cfg.vocab_size, shape=c["shape"], seed=c["batch_seed"])
    o = O.tvnet2_forward(params, cfg, batch, alpha=0.1, beta=0.5, **g["flags"])
    _close(o["loss"], g["loss"], rtol=1e-5, atol=1e-5)
    assert o["logits"] == g["logits"]
    if g["prob_loss"] is not None:
        _close(o["prob_loss"], g["prob_loss"], rtol=1e-5, atol=1e-3)
    if g["img_loss"] is not None:
        _close(o["img_loss"], g["img_loss"], rtol=1e-5, atol=1e-6)
    o["loss"].backward()
    fp = grad_fingerprint([(k, v.grad) for k, v in params.items()])
    _check_fp(fp, g["grad_fp"], 5e-4)


def test_tvnet_span_matches_reference_golden(golden_dir):
    """Span variant TVNetSAModel: oracle restatement vs the unmodified reference."""
    g = _load(golden_dir, "tvnet_span_roberta")
    c = CASES["tvnet_span_roberta"]
    cfg = ocfg_for(c["kind"])
    params = S.init_params(cfg, seed=c["param_seed"], ln_jitter=0.05, with_span=True)
    params = {k: v.requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}
    batch = S.make_span_batch(c["B"], c["L"], M=c["M"], vocab=cfg.vocab_size, shape=c["shape"], seed=c["batch_seed"])
    o = O.tvnet_forward(params, cfg, batch, beta=0.5, num_epochs=30)
    _close(o["loss"], g["loss"], rtol=1e-5, atol=1e-5)
    _close(o["tot_loss"], g["tot_loss"], rtol=1e-5, atol=1e-5)
    _close(o["prob_loss"], g["prob_loss"], rtol=1e-5, atol=1e-2)
    _close(o["logits"], g["logits"], rtol=1e-4, atol=1e-5)
    o["loss"].backward()
    fp = grad_fingerprint([(k, v.grad) for k, v in params.items() if k in g["grad_fp"]])
    _check_fp(fp, {k: v for k, v in g["grad_fp"].items() if k in fp}, 5e-4)
    for k in ("dense.weight", "unary_affine.weight", "binary_affine.weight", "classifier.weight"):
        assert k in fp and g["grad_fp"][k] is not None, k


def test_crf_against_bruteforce():
    """pytorch-crf is absent (parity unpinned): check the restated forward algorithm / Viterbi
    against explicit enumeration of all tag paths on a tiny problem."""
    import itertools
    torch.manual_seed(3)
    B, L, T = 3, 4, 3
    em = torch.randn(B, L, T)
    start, end, trans = torch.randn(T), torch.randn(T), torch.randn(T, T)
    mask = torch.tensor([[1, 1, 1, 1], [1, 1, 0, 0], [1, 0, 0, 0]])
    tags = torch.randint(0, T, (B, L))
    llh = O.crf_log_likelihood(em, tags, mask, start, end, trans)
    dec = O.crf_decode(em, mask, start, end, trans)
    for b in range(B):
        n = int(mask[b].sum())

        def score(path):
            s = start[path[0]] + em[b, 0, path[0]]
            for i in range(1, n):
                s = s + trans[path[i - 1], path[i]] + em[b, i, path[i]]
            return s + end[path[n - 1]]
        allp = list(itertools.product(range(T), repeat=n))
        scores = torch.stack([score(pth) for pth in allp])
        logz = torch.logsumexp(scores, 0)
        ref = score(tags[b, :n].tolist()) - logz
        torch.testing.assert_close(llh[b], ref, rtol=1e-5, atol=1e-5)
        assert list(allp[int(scores.argmax())]) == dec[b]
