"""Synthetic Twitter2015/2017-shaped batches and random-init parameters.

There is no network in the box, so no datasets or checkpoints: shapes and id/label conventions follow
the reference's feature builder (ids padded with 0, `modules/dataset.py:414-415`; labels 0=pad,
1..8 interior, 9=[CLS], 10=[SEP], `modules/dataset.py:211-212,358`), generated on the CPU generator so
the CPU oracle and the GPU path see identical bytes.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

LENGTH_STATS = {           # (mean, std, min) of the number of real tokens; stand-ins (none published)
    "twitter2015": (30.0, 10.0, 8),
    "twitter2017": (34.0, 12.0, 8),
    "longaux": (200.0, 40.0, 64),
}


def make_batch(B: int, L: int, *, vocab: int = 50265, shape: str = "twitter2017", n_aux: int = 3,
               n_anp: int = 2089, seed: int = 2024, with_images: bool = True) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    mean, std, lo = LENGTH_STATS[shape]
    lens = torch.clamp(torch.round(torch.randn(B, generator=g) * std + mean), lo, L).long()
    ar = torch.arange(L).unsqueeze(0)
    mask = (ar < lens.unsqueeze(1)).long()
    ids = torch.randint(3, vocab, (B, L), generator=g) * mask                 # pad id 0
    labels = torch.randint(1, 9, (B, L), generator=g) * mask
    labels[:, 0] = 9
    labels[torch.arange(B), lens - 1] = 10
    batch = {"input_ids": ids, "attention_mask": mask, "token_type_ids": torch.zeros(B, L, dtype=torch.long),
             "labels": labels}
    if with_images:
        batch["images"] = torch.randn(B, 3840, 2, 2, generator=g).abs()
        batch["aux_imgs"] = torch.randn(B, n_aux, 3840, 2, 2, generator=g).abs()
        batch["imagelabel"] = torch.softmax(torch.randn(B, n_anp, generator=g), dim=-1)
    return batch


def make_prefix(B: int, n_layers: int, n_heads: int, P: int, d: int = 64, seed: int = 7):
    """Prefix K/V at the attention boundary: "R object regions" == P=R prefix rows per layer."""
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(B, n_heads, P, d, generator=g), torch.randn(B, n_heads, P, d, generator=g))
            for _ in range(n_layers)]


def encoder_param_shapes(cfg, prefix="bert.") -> "OrderedDict[str, tuple]":
    """state_dict keys and shapes of the reference encoder, in registration order
    (models/modeling_roberta.py:76-100,174-176,291-292,358,374-375,672)."""
    H, I = cfg.hidden_size, cfg.intermediate_size
    s = OrderedDict()
    e = prefix + "embeddings."
    s[e + "word_embeddings.weight"] = (cfg.vocab_size, H)
    s[e + "position_embeddings.weight"] = (cfg.max_position_embeddings, H)
    s[e + "token_type_embeddings.weight"] = (cfg.type_vocab_size, H)
    s[e + "LayerNorm.weight"] = (H,)
    s[e + "LayerNorm.bias"] = (H,)
    for i in range(cfg.num_hidden_layers):
        l = f"{prefix}encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            s[l + f"attention.self.{nm}.weight"] = (H, H)
            s[l + f"attention.self.{nm}.bias"] = (H,)
        s[l + "attention.output.dense.weight"] = (H, H)
        s[l + "attention.output.dense.bias"] = (H,)
        s[l + "attention.output.LayerNorm.weight"] = (H,)
        s[l + "attention.output.LayerNorm.bias"] = (H,)
        s[l + "intermediate.dense.weight"] = (I, H)
        s[l + "intermediate.dense.bias"] = (I,)
        s[l + "output.dense.weight"] = (H, I)
        s[l + "output.dense.bias"] = (H,)
        s[l + "output.LayerNorm.weight"] = (H,)
        s[l + "output.LayerNorm.bias"] = (H,)
    s[prefix + "pooler.dense.weight"] = (H, H)
    s[prefix + "pooler.dense.bias"] = (H,)
    return s


def init_params(cfg, *, seed: int = 1234, with_fusion: bool = True, n_aux: int = 3, n_anp: int = 2089,
                num_labels: int = 11, probe_rank: int = None, ln_jitter: float = 0.0,
                with_span: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Random-init parameter dict keyed like the reference state_dict.

    Encoder: the reference's `_init_weights` (models/modeling_roberta.py:695-709): Linear/Embedding
    N(0,0.02), biases 0, pad rows zero, LayerNorm (1,0).  `ln_jitter`>0 perturbs biases and LayerNorm
    affine parameters so parity tests exercise them.  Heads use small normals (nn.Linear default
    init is not part of the path's arithmetic)."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    H = cfg.hidden_size
    for k, shp in encoder_param_shapes(cfg).items():
        if k.endswith("LayerNorm.weight"):
            t = torch.ones(shp) + ln_jitter * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            t = ln_jitter * torch.randn(shp, generator=g)
        else:
            t = torch.randn(shp, generator=g) * 0.02
        p[k] = t
    p["bert.embeddings.word_embeddings.weight"][cfg.pad_token_id].zero_()
    if cfg.kind == "roberta":
        p["bert.embeddings.position_embeddings.weight"][cfg.pad_token_id].zero_()
    if with_fusion:
        def lin(name, out_f, in_f, std=None):
            std = std if std is not None else (1.0 / in_f) ** 0.5
            p[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * std
            p[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02
        lin("encoder_conv.0", 800, 3840, 0.02)
        lin("encoder_conv.2", 4 * 2 * H, 800)
        for i in range(cfg.num_hidden_layers):
            lin(f"projectors.{i}", 4, 4 * 2 * H)
        lin("img_classifier", n_anp, 4 * 2 * H)
        for k in range(n_aux):
            lin(f"aux_img_classifier.{k}", n_anp, 4 * 2 * H)
    p["crf.start_transitions"] = torch.rand(num_labels, generator=g) * 0.2 - 0.1
    p["crf.end_transitions"] = torch.rand(num_labels, generator=g) * 0.2 - 0.1
    p["crf.transitions"] = torch.rand(num_labels, num_labels, generator=g) * 0.2 - 0.1
    p["fc.weight"] = torch.randn(num_labels, H, generator=g) * (1.0 / H) ** 0.5
    p["fc.bias"] = torch.randn(num_labels, generator=g) * 0.02
    r = probe_rank if probe_rank is not None else H // 2
    p["oneWordpsdProbe.oneWordpsdProbe.proj"] = torch.rand(H, r, generator=g) * 0.1 - 0.05   # probes/probe.py:60
    if with_span:
        # heads of the span variant TVNetSAModel (models/bert_model.py:205-212)
        g2 = torch.Generator().manual_seed(seed + 7919)
        for name, out_f in (("dense", H), ("unary_affine", 1), ("binary_affine", 2), ("classifier", 4)):
            p[name + ".weight"] = torch.randn(out_f, H, generator=g2) * (1.0 / H) ** 0.5
            p[name + ".bias"] = torch.randn(out_f, generator=g2) * 0.02
    return p


def make_span_batch(B: int, L: int, *, M: int = 20, vocab: int = 50265, shape: str = "twitter2015", n_aux: int = 3,
                    seed: int = 2024, with_images: bool = True) -> Dict[str, torch.Tensor]:
    """Synthetic batch of the span variant (TVNetSAModel.forward, models/bert_model.py:246-252): multi-hot
    start/end positions, up to M candidate spans per sentence (padded with the (0, 0) span, label mask 0),
    4-way polarity labels."""
    b = make_batch(B, L, vocab=vocab, shape=shape, n_aux=n_aux, seed=seed, with_images=with_images)
    b.pop("labels")
    b.pop("imagelabel", None)
    g = torch.Generator().manual_seed(seed + 31337)
    lens = b["attention_mask"].sum(1)
    starts = torch.zeros(B, M, dtype=torch.long)
    ends = torch.zeros(B, M, dtype=torch.long)
    lmask = torch.zeros(B, M, dtype=torch.long)
    sp = torch.zeros(B, L, dtype=torch.long)
    ep = torch.zeros(B, L, dtype=torch.long)
    for i in range(B):
        n = int(torch.randint(1, min(M, 6) + 1, (1,), generator=g))
        for j in range(n):
            s0 = int(torch.randint(1, max(2, int(lens[i]) - 1), (1,), generator=g))
            w = int(torch.randint(0, 4, (1,), generator=g))
            e0 = min(s0 + w, int(lens[i]) - 1)
            starts[i, j], ends[i, j], lmask[i, j] = s0, e0, 1
            sp[i, s0] = 1
            ep[i, e0] = 1
    b.update(start_positions=sp, end_positions=ep, span_starts=starts, span_ends=ends,
             polarity_labels=torch.randint(0, 4, (B, M), generator=g) * lmask, label_masks=lmask)
    return b
