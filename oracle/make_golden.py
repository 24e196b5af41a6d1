"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by running the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):  python -m oracle.make_golden
Inputs and weights are regenerated from seeds (mtvaf_b200.synthetic), so the fixtures hold only the
reference's OUTPUTS (kept small: slices + checksums).  The GPU box has no /root/reference; tests
there compare the CUDA path and the oracle with these files.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim                      # noqa: E402
from oracle import mtvaf_oracle as O             # noqa: E402
from mtvaf_b200 import synthetic as S            # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (kind, B, L, shape, batch_seed, param_seed)
    "tvnet2_roberta": dict(kind="roberta", B=4, L=32, shape="twitter2015", batch_seed=11, param_seed=101),
    "encoder_roberta_p36": dict(kind="roberta", B=2, L=48, shape="twitter2017", batch_seed=12, param_seed=102, P=36),
    "encoder_bert": dict(kind="bert", B=2, L=40, shape="twitter2015", batch_seed=13, param_seed=103),
    "tvnet_span_roberta": dict(kind="roberta", B=4, L=32, shape="twitter2015", batch_seed=14, param_seed=104, M=6),
    "tvnet2_bert": dict(kind="bert", B=3, L=24, shape="twitter2017", batch_seed=15, param_seed=105),
}


def hf_config(ocfg):
    from transformers import RobertaConfig, BertConfig
    cls = RobertaConfig if ocfg.kind == "roberta" else BertConfig
    return cls(vocab_size=ocfg.vocab_size, hidden_size=ocfg.hidden_size,
               num_hidden_layers=ocfg.num_hidden_layers, num_attention_heads=ocfg.num_attention_heads,
               intermediate_size=ocfg.intermediate_size, max_position_embeddings=ocfg.max_position_embeddings,
               type_vocab_size=ocfg.type_vocab_size, layer_norm_eps=ocfg.layer_norm_eps,
               pad_token_id=ocfg.pad_token_id, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)


def ocfg_for(kind):
    return O.EncoderCfg.roberta_base() if kind == "roberta" else O.EncoderCfg.bert_base()


def grad_fingerprint(named_grads):
    """Small per-parameter fingerprint: L2 norm, sum, and the first 4 elements."""
    fp = {}
    for k, g in named_grads:
        if g is None:
            fp[k] = None
            continue
        g = g.detach().double().flatten()
        fp[k] = torch.cat([g.norm().view(1), g.sum().view(1), g[:4]]).float()
    return fp


def gen_tvnet2(name, c):
    ocfg = ocfg_for(c["kind"])
    params = S.init_params(ocfg, seed=c["param_seed"], ln_jitter=0.05)
    batch = S.make_batch(c["B"], c["L"], vocab=ocfg.vocab_size, shape=c["shape"], seed=c["batch_seed"])
    # the reference picks the backbone by name: "roberta" in args.bert_name (models/bert_model.py:425-429)
    args = ref_shim.make_args(**({"bert_name": "bert-base-uncased"} if c["kind"] == "bert" else {}))
    model = ref_shim.build_reference_tvnet2(hf_config(ocfg), args, list(range(10)))
    missing = model.load_state_dict(params, strict=False)
    assert not missing.unexpected_keys, missing
    # only non-persistent / index buffers may be missing
    assert all("position_ids" in k or "token_type_ids" in k for k in missing.missing_keys), missing
    model.eval()
    out, prob_loss, img_loss = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                                     token_type_ids=batch["token_type_ids"], labels=batch["labels"],
                                     imagelabel=batch["imagelabel"], images=batch["images"],
                                     aux_imgs=batch["aux_imgs"])
    out.loss.backward()
    with torch.no_grad():
        pkv, _, _ = model.get_visual_prompt(batch["images"], batch["aux_imgs"], batch["imagelabel"])
        P = pkv[0][0].shape[2]
        full_mask = torch.cat([torch.ones(c["B"], P), batch["attention_mask"].float()], dim=1)
        enc = model.bert(input_ids=batch["input_ids"], attention_mask=full_mask,
                         token_type_ids=batch["token_type_ids"], past_key_values=pkv,
                         output_attentions=False, output_hidden_states=True, return_dict=True)
        hs = enc["hidden_states"]
        emissions = model.fc(enc["last_hidden_state"])
        norms = model.oneWordpsdProbe.oneWordpsdProbe(hs[7])
        labels = model.oneWordpsdProbe.constructLabel(norms)
    gold = {
        "case": c, "loss": out.loss.detach(), "prob_loss": prob_loss.detach(), "img_loss": img_loss.detach(),
        "logits": out.logits, "emissions": emissions, "norms": norms, "pseudo_labels": labels,
        "hs7_slice": hs[7][:, :, :16].clone(), "last_slice": hs[12][:, :, :16].clone(),
        "hs_norms": torch.stack([h.double().norm() for h in hs]).float(),
        "prefix_k0": pkv[0][0].clone(), "prefix_v11_slice": pkv[11][1][:, :, :, :8].clone(),
        "grad_fp": grad_fingerprint([(k, v.grad) for k, v in model.named_parameters()]),
    }
    torch.save(gold, os.path.join(GOLD, name + ".pt"))
    print(name, "loss", float(out.loss), "prob", float(prob_loss), "img", float(img_loss))


def gen_tvnet_span(name, c):
    """Span variant TVNetSAModel (models/bert_model.py:192-414), eval mode, with prefix and probe."""
    ocfg = ocfg_for(c["kind"])
    params = S.init_params(ocfg, seed=c["param_seed"], ln_jitter=0.05, with_span=True)
    batch = S.make_span_batch(c["B"], c["L"], M=c["M"], vocab=ocfg.vocab_size, shape=c["shape"], seed=c["batch_seed"])
    args = ref_shim.make_args(vao=False)
    model = ref_shim.build_reference_tvnet2(hf_config(ocfg), args, list(range(10)), cls_name="TVNetSAModel")
    missing = model.load_state_dict({k: v for k, v in params.items() if k in model.state_dict()}, strict=False)
    assert all("position_ids" in k or "token_type_ids" in k for k in missing.missing_keys), missing
    model.eval()
    out, prob_loss, tot_loss = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                                     token_type_ids=batch["token_type_ids"],
                                     start_positions=batch["start_positions"], end_positions=batch["end_positions"],
                                     span_starts=batch["span_starts"], span_ends=batch["span_ends"],
                                     polarity_labels=batch["polarity_labels"], label_masks=batch["label_masks"],
                                     images=batch["images"], aux_imgs=batch["aux_imgs"])
    out.loss.backward()
    gold = {"case": c, "loss": out.loss.detach(), "prob_loss": prob_loss.detach(), "tot_loss": tot_loss.detach(),
            "logits": out.logits.detach().clone(),
            "grad_fp": grad_fingerprint([(k, v.grad) for k, v in model.named_parameters()])}
    torch.save(gold, os.path.join(GOLD, name + ".pt"))
    print(name, "loss", float(out.loss), "prob", float(prob_loss), "tot", float(tot_loss))


def gen_encoder(name, c):
    R = ref_shim.load_reference()
    ocfg = ocfg_for(c["kind"])
    params = S.init_params(ocfg, seed=c["param_seed"], ln_jitter=0.05, with_fusion=False)
    batch = S.make_batch(c["B"], c["L"], vocab=ocfg.vocab_size, shape=c["shape"], seed=c["batch_seed"],
                         with_images=False)
    cls = R.roberta.RobertaModel if c["kind"] == "roberta" else R.bert.BertModel
    torch.manual_seed(0)
    model = cls(hf_config(ocfg))
    sd = {k[len("bert."):]: v for k, v in params.items() if k.startswith("bert.")}
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    model.eval()
    P = c.get("P", 0)
    pkv = None
    mask = batch["attention_mask"].float()
    if P:
        pkv = S.make_prefix(c["B"], ocfg.num_hidden_layers, ocfg.num_attention_heads, P,
                            ocfg.hidden_size // ocfg.num_attention_heads, seed=c["batch_seed"] + 1000)
        pkv = [(k.requires_grad_(), v.requires_grad_()) for k, v in pkv]
        mask = torch.cat([torch.ones(c["B"], P), mask], dim=1)
    tt = batch["token_type_ids"]
    if c["kind"] == "bert":
        tt = (torch.arange(c["L"]).unsqueeze(0).expand(c["B"], -1) % 2).contiguous()
    enc = model(input_ids=batch["input_ids"], attention_mask=mask, token_type_ids=tt,
                past_key_values=pkv, output_attentions=True, output_hidden_states=True, return_dict=True)
    hs = enc["hidden_states"]
    # scalar objective touching hidden states 7 and 12 so gradients flow like the real loss
    g = torch.Generator().manual_seed(99)
    w7 = torch.randn(hs[7].shape, generator=g)
    w12 = torch.randn(hs[12].shape, generator=g)
    obj = (hs[7] * w7).sum() + (hs[12] * w12).sum()
    obj.backward()
    gold = {
        "case": c, "hs_norms": torch.stack([h.double().norm() for h in hs]).float(),
        "last": hs[-1].detach().clone(), "hs7_slice": hs[7][:, :, :16].detach().clone(),
        "emb_slice": hs[0][:, :, :16].detach().clone(),
        "pooler": enc["pooler_output"].detach().clone(),
        "attn0_slice": enc["attentions"][0][:, :2, :8, :].detach().clone(),
        "obj": obj.detach(),
        "grad_fp": grad_fingerprint([("bert." + k, v.grad) for k, v in model.named_parameters()]),
    }
    if pkv is not None:
        gold["dk0"] = pkv[0][0].grad.clone()
        gold["dv11_slice"] = pkv[11][1].grad[:, :, :, :8].clone()
    torch.save(gold, os.path.join(GOLD, name + ".pt"))
    print(name, "obj", float(obj))


def gen_probe():
    """Known-answer vectors with the SHIPPED layer-7/4 probe matrices."""
    R = ref_shim.load_reference()
    gold = {}
    for lvl in (4, 7):
        path = os.path.join(ref_shim.REF_ROOT, "probes", f"psdProbe_base_savel{lvl}.pt")
        mod = torch.load(path, map_location="cpu", weights_only=False)
        proj = mod.state_dict()["oneWordpsdProbe.proj"]
        torch.manual_seed(1234)
        x = torch.randn(2, 6, 768)
        one = R.probe.OneWordPSDProbe({"probe": {"maximum_rank": 384}, "model": {"hidden_dim": 768}})
        with torch.no_grad():
            one.proj.copy_(proj)
            norms = one(x)
            labels = R.label.ConstructLabelGaget(None)(norms)
            two = R.probe.TwoWordPSDProbe({"probe": {"maximum_rank": 384}, "model": {"hidden_dim": 768},
                                           "device": "cpu"})
            two.proj.copy_(proj)
            dist = two(x)
            mod.eval()
            loss = mod(x)
        import hashlib
        gold[f"l{lvl}"] = {"norms": norms, "labels": labels, "dist": dist, "loss": loss,
                           "proj_fro": proj.norm(), "proj_sha256": hashlib.sha256(proj.numpy().tobytes()).hexdigest(),
                           "proj_colsum": proj.sum(0), "proj_rowsum": proj.sum(1)}
        print("probe l%d" % lvl, norms[0].tolist(), labels[0].tolist(), float(loss), float(dist[0, 0, 1]))
    # ConstructLabelGaget known-answer
    v = torch.tensor([[7.0, 1.1, 0.2, 2.9, 1.3, 2.6], [2.5, 2.5, 0.0, 9.0, 3.5, 2.5]])
    gold["label_kat_in"] = v
    gold["label_kat_out"] = R.label.ConstructLabelGaget(None)(v)
    g = torch.Generator().manual_seed(5)
    big = torch.rand(6, 128, generator=g) * 40
    big[1, 5] = big[1, 77]           # exact ties exercise sort stability
    big[2] = torch.round(big[2] * 2) / 2   # exact .5 values exercise the tie rule
    gold["label_big_in"] = big
    gold["label_big_out"] = R.label.ConstructLabelGaget(None)(big)
    torch.save(gold, os.path.join(GOLD, "probe_kat.pt"))


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    gen_probe()
    gen_encoder("encoder_roberta_p36", CASES["encoder_roberta_p36"])
    gen_encoder("encoder_bert", CASES["encoder_bert"])
    gen_tvnet2("tvnet2_roberta", CASES["tvnet2_roberta"])
    gen_tvnet2("tvnet2_bert", CASES["tvnet2_bert"])
    gen_tvnet_span("tvnet_span_roberta", CASES["tvnet_span_roberta"])


if __name__ == "__main__":
    main()
