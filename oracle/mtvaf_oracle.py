"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of MTVAF's data-parallel hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module, and only as the checker / reported CPU baseline.  The product package
`mtvaf_b200/` never imports it and has no CPU fallback.

The path is floating point (fp32 in the reference), so the restatement is written with plain
torch-CPU tensor algebra in functional form (every function is differentiable, so autograd on the
oracle gives the reference gradients); integer/index work (position ids, K/V reshape, pseudo labels,
Viterbi) is bit-exact.  Every function cites the reference file:line it restates
(paths relative to the reference root).

Parity pinning:
  * pinned against the reference ITSELF run in the authoring container through `oracle/ref_shim.py`
    (`tests/test_oracle_vs_reference.py`, and the committed `tests/golden/*.pt` produced by
    `oracle/make_golden.py`), plus the survey's known-answer vectors for ConstructLabelGaget and the
    shipped layer-7 probe matrix;
  * the linear-chain CRF (`pytorch-crf`, imported by models/bert_model.py:7 but absent from the
    reference tree and from this image, no version pinned; lineage suggests 0.7.2) is restated from
    its published algorithm -> for `crf_*` **parity is unpinned**; it is cross-checked against a
    brute-force enumeration in tests instead.

Parameters are passed as a flat dict keyed by the reference's own state_dict names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
class EncoderCfg:
    """The few config fields the path reads (transformers RobertaConfig/BertConfig names)."""

    def __init__(self, kind="roberta", vocab_size=50265, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, max_position_embeddings=514,
                 type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1):
        self.kind = kind
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.layer_norm_eps = layer_norm_eps
        self.pad_token_id = pad_token_id

    @staticmethod
    def roberta_base(**kw):
        return EncoderCfg(**kw)

    @staticmethod
    def bert_base(**kw):
        d = dict(kind="bert", vocab_size=30522, max_position_embeddings=512, type_vocab_size=2,
                 layer_norm_eps=1e-12, pad_token_id=0)
        d.update(kw)
        return EncoderCfg(**d)

    @staticmethod
    def roberta_large(**kw):
        d = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
        d.update(kw)
        return EncoderCfg(**d)


# --------------------------------------------------------------------------------------------
# encoder (models/modeling_roberta.py, models/modeling_bert.py)
# --------------------------------------------------------------------------------------------
def roberta_position_ids(input_ids: Tensor, padding_idx: int = 1) -> Tensor:
    """models/modeling_roberta.py:1706-1719 with past_key_values_length forced to 0 (:911).
    int64 in, int64 out, bit-exact.  NB the datasets pad ids with 0 while padding_idx is 1, so padded
    positions keep counting."""
    mask = input_ids.ne(padding_idx).to(torch.int32)
    inc = torch.cumsum(mask, dim=1).to(torch.int32) * mask
    return inc.long() + padding_idx


def embeddings(p: Dict[str, Tensor], cfg: EncoderCfg, input_ids: Tensor, token_type_ids: Tensor,
               prefix: str = "bert.") -> Tuple[Tensor, Tensor]:
    """RobertaEmbeddings.forward models/modeling_roberta.py:102-140; BertEmbeddings.forward
    models/modeling_bert.py:188-222 (absolute position slice :199).  Dropout is identity (eval)."""
    B, L = input_ids.shape
    if cfg.kind == "roberta":
        pos = roberta_position_ids(input_ids, cfg.pad_token_id)
    else:
        pos = torch.arange(L, dtype=torch.long).unsqueeze(0).expand(B, L)
    # nn.Embedding(padding_idx=...) (:78,98-100; BERT models/modeling_bert.py:170): the pad row gets NO
    # gradient.  RoBERTa: word and position tables both use padding_idx=pad_token_id; BERT: word only.
    pad = cfg.pad_token_id
    e = (F.embedding(input_ids, p[prefix + "embeddings.word_embeddings.weight"], padding_idx=pad)
         + F.embedding(token_type_ids, p[prefix + "embeddings.token_type_embeddings.weight"])
         + F.embedding(pos, p[prefix + "embeddings.position_embeddings.weight"],
                       padding_idx=pad if cfg.kind == "roberta" else None))
    e = F.layer_norm(e, (cfg.hidden_size,), p[prefix + "embeddings.LayerNorm.weight"],
                     p[prefix + "embeddings.LayerNorm.bias"], cfg.layer_norm_eps)
    return e, pos


def extended_mask(attention_mask: Tensor) -> Tensor:
    """HF 4.x get_extended_attention_mask as called at models/modeling_roberta.py:926 and restated
    by the reference itself at :997-1000: (1 - m)[:,None,None,:] * -10000.0."""
    return (1.0 - attention_mask[:, None, None, :].to(torch.float32)) * -10000.0


def gelu_erf(x: Tensor) -> Tensor:
    """ACT2FN['gelu'] (models/modeling_roberta.py:26,360): exact erf form."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def self_attention(p, cfg, lp: str, x: Tensor, ext_mask: Tensor,
                   pkv: Optional[Tuple[Tensor, Tensor]], return_probs=False):
    """RobertaSelfAttention.forward models/modeling_roberta.py:191-284 (BERT :255-342):
    prefix K/V are concatenated IN FRONT of the text K/V (:221-222), one softmax over P+L keys."""
    B, L, H = x.shape
    nh = cfg.num_attention_heads
    d = H // nh

    def heads(t):  # transpose_for_scores :186-189
        return t.view(B, -1, nh, d).permute(0, 2, 1, 3)

    q = heads(F.linear(x, p[lp + "attention.self.query.weight"], p[lp + "attention.self.query.bias"]))
    k = heads(F.linear(x, p[lp + "attention.self.key.weight"], p[lp + "attention.self.key.bias"]))
    v = heads(F.linear(x, p[lp + "attention.self.value.weight"], p[lp + "attention.self.value.bias"]))
    if pkv is not None:
        k = torch.cat([pkv[0], k], dim=2)
        v = torch.cat([pkv[1], v], dim=2)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)          # :240,258
    s = s + ext_mask                                                 # :261
    pr = torch.softmax(s, dim=-1)                                    # :264
    ctx = torch.matmul(pr, v).permute(0, 2, 1, 3).contiguous().view(B, L, H)   # :274-278
    return (ctx, pr) if return_probs else ctx


def encoder_layer(p, cfg, lp: str, x: Tensor, ext_mask: Tensor, pkv):
    """RobertaLayer.forward models/modeling_roberta.py:400-469 = attention (:413) ->
    RobertaSelfOutput (:295-299) -> RobertaIntermediate (:364-367) -> RobertaOutput (:378-382)."""
    H = cfg.hidden_size
    ctx = self_attention(p, cfg, lp, x, ext_mask, pkv)
    a = F.linear(ctx, p[lp + "attention.output.dense.weight"], p[lp + "attention.output.dense.bias"])
    a = F.layer_norm(a + x, (H,), p[lp + "attention.output.LayerNorm.weight"],
                     p[lp + "attention.output.LayerNorm.bias"], cfg.layer_norm_eps)
    h = gelu_erf(F.linear(a, p[lp + "intermediate.dense.weight"], p[lp + "intermediate.dense.bias"]))
    o = F.linear(h, p[lp + "output.dense.weight"], p[lp + "output.dense.bias"])
    o = F.layer_norm(o + a, (H,), p[lp + "output.LayerNorm.weight"], p[lp + "output.LayerNorm.bias"],
                     cfg.layer_norm_eps)
    return o


def encoder_forward(p, cfg: EncoderCfg, input_ids: Tensor, attention_mask: Tensor,
                    token_type_ids: Tensor, past_key_values=None, prefix: str = "bert."):
    """RobertaModel.forward models/modeling_roberta.py:850-978 / RobertaEncoder.forward :480-568.
    `attention_mask` is [B, P+L] (prefix columns first, models/bert_model.py:490-492).
    Returns dict(last_hidden_state, hidden_states (n+1 tuple, [0] = embeddings), position_ids)."""
    x, pos = embeddings(p, cfg, input_ids, token_type_ids, prefix)
    ext = extended_mask(attention_mask)
    hs = [x]
    for i in range(cfg.num_hidden_layers):
        pkv = past_key_values[i] if past_key_values is not None else None
        x = encoder_layer(p, cfg, f"{prefix}encoder.layer.{i}.", x, ext, pkv)
        hs.append(x)
    return {"last_hidden_state": x, "hidden_states": tuple(hs), "position_ids": pos}


def pooler(p, x: Tensor, prefix="bert."):
    """RobertaPooler.forward models/modeling_roberta.py:675-681 (dead on the TVNetSAModel2 path)."""
    return torch.tanh(F.linear(x[:, 0], p[prefix + "pooler.dense.weight"], p[prefix + "pooler.dense.bias"]))


# --------------------------------------------------------------------------------------------
# visual prompt / fusion stack (models/bert_model.py:534-588, twin :379-414)
# --------------------------------------------------------------------------------------------
def visual_prompt(p, images: Tensor, aux_imgs: Tensor, imagelabel: Optional[Tensor],
                  vao: bool = True, n_layers: int = 12, hidden: int = 768, n_heads: int = 12):
    """TVNetSAModel2.get_visual_prompt models/bert_model.py:534-588 at the fusion boundary.

    images  [B,3840,2,2]  (the channel-concatenated 4-level pyramid, :538)
    aux_imgs[B,n_aux,3840,2,2]
    Returns (list of n_layers (K,V) each [B,n_heads,4*(1+n_aux),hidden/n_heads], img_tag_loss,
    [aux_img_tag_loss...]).  Dropout (img_dropout, :551) is identity (eval).
    """
    B = images.shape[0]
    n_aux = aux_imgs.shape[1]
    feats = [images.reshape(B, 4, -1)] + [aux_imgs[:, i].reshape(B, 4, -1) for i in range(n_aux)]  # :538-539

    def mlp(t):   # encoder_conv :446-454,541-542
        t = torch.tanh(F.linear(t, p["encoder_conv.0.weight"], p["encoder_conv.0.bias"]))
        return F.linear(t, p["encoder_conv.2.weight"], p["encoder_conv.2.bias"])

    guids = [mlp(f) for f in feats]                      # each [B,4,4*2*hidden]
    splits = [g.split(hidden * 2, dim=-1) for g in guids]  # :544-545, 4 x [B,4,2*hidden]

    img_loss = torch.zeros(())
    aux_losses: List[Tensor] = []
    if vao:                                               # :549-563
        names = ["img_classifier"] + [f"aux_img_classifier.{k}" for k in range(n_aux)]
        for j, (g, nm) in enumerate(zip(guids, names)):
            logits = F.linear(g.mean(dim=1), p[nm + ".weight"], p[nm + ".bias"])
            sm = torch.softmax(logits, dim=-1)
            loss = F.kl_div(sm.log(), imagelabel, reduction="batchmean")
            if j == 0:
                img_loss = loss
            else:
                aux_losses.append(loss)

    result = []
    for idx in range(n_layers):                           # :566-587
        W, b = p[f"projectors.{idx}.weight"], p[f"projectors.{idx}.bias"]
        kvs = []
        for sp in splits:
            s = torch.stack(sp).sum(0).view(B, -1) / 4    # :567,576  [B,4*2*hidden]
            g = torch.softmax(F.leaky_relu(F.linear(s, W, b)), dim=-1)   # [B,4]
            kv = torch.zeros_like(sp[0])
            for i in range(4):
                kv = kv + g[:, i].view(-1, 1, 1) * sp[i]  # einsum('bg,blh->blh') :572,580
            kvs.append(kv)
        kv = torch.cat(kvs, dim=1)                        # [B,4*(1+n_aux),2*hidden] :583
        k, v = kv.split(hidden, dim=-1)                   # :584
        # NB plain reshape, NOT a head transpose (:585): mixes sequence and head dims.
        k = k.reshape(B, n_heads, -1, hidden // n_heads).contiguous()
        v = v.reshape(B, n_heads, -1, hidden // n_heads).contiguous()
        result.append((k, v))
    return result, img_loss, aux_losses


# --------------------------------------------------------------------------------------------
# psdProbe (probes/)
# --------------------------------------------------------------------------------------------
def one_word_psd_probe(x: Tensor, proj: Tensor) -> Tensor:
    """OneWordPSDProbe.forward probes/probe.py:62-79: squared L2 norm of x @ proj per token."""
    t = torch.matmul(x, proj)
    return (t * t).sum(-1)


def two_word_psd_probe(x: Tensor, proj: Tensor) -> Tensor:
    """TwoWordPSDProbe.forward probes/probe.py:25-46 (defined, never called by the reference):
    explicit broadcast differences, so the diagonal is exactly 0 and D is exactly symmetric."""
    t = torch.matmul(x, proj)
    diffs = t.unsqueeze(2) - t.unsqueeze(1)
    return (diffs * diffs).sum(-1)


def construct_label(norms: Tensor) -> Tensor:
    """ConstructLabelGaget.forward probes/constructLabel.py:11-29, bit-exact.

    Per row: STABLE ascending sort of the fp32 norms (Python list.sort on 0-d tensors); rank0 -> 1,
    rank1 -> 2, then lab_j = lab_{j-1} if |v_j - lab_{j-1}| < |lab_{j-1} + 1 - v_j| else lab_{j-1}+1
    with both sides evaluated in fp32; labels scattered back to the original positions.  Padding
    positions are included (no mask).  Output fp32 integer-valued, no grad."""
    import numpy as np
    v = norms.detach().to(torch.float32).cpu().numpy()
    B, L = v.shape
    out = np.zeros((B, L), dtype=np.float32)
    for i in range(B):
        order = np.argsort(v[i], kind="stable")
        lab = np.float32(0)
        for r, j in enumerate(order):
            if r == 0:
                lab = np.float32(1)
            elif r == 1:
                lab = np.float32(2)
            else:
                x = v[i, j]
                if not (np.abs(np.float32(x - lab)) < np.abs(np.float32(np.float32(lab + np.float32(1)) - x))):
                    lab = np.float32(lab + np.float32(1))
            out[i, j] = lab
    return torch.from_numpy(out)


def probe_loss(x: Tensor, proj: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """probe.forward probes/probe_trainModel.py:15-26: MSE(norms, pseudo labels), labels constant."""
    norms = one_word_psd_probe(x, proj)
    labels = construct_label(norms)
    return F.mse_loss(norms, labels), norms, labels


def combine_loss(loss: Tensor, prob_loss: Tensor, beta: float, epoch: int = 30) -> Tensor:
    """CombineLoss.forward probes/loss.py:13-18 (epoch is the constant 30, models/bert_model.py:523)."""
    if prob_loss.item() > 0.1:
        return loss + prob_loss * torch.tensor(beta) * torch.tensor(pow(2, -epoch))
    return loss


# --------------------------------------------------------------------------------------------
# linear-chain CRF (pytorch-crf semantics; call sites models/bert_model.py:464,511,521)
# --------------------------------------------------------------------------------------------
def crf_log_likelihood(emissions: Tensor, tags: Tensor, mask: Tensor, start: Tensor, end: Tensor,
                       trans: Tensor) -> Tensor:
    """Per-sequence log-likelihood  score(tags) - logZ  (batch_first=True; mask[:,0] must be on).
    emissions [B,L,T] float, tags [B,L] int64, mask [B,L] {0,1}.  Returns [B]."""
    B, L, T = emissions.shape
    m = mask.to(emissions.dtype)
    ar = torch.arange(B)
    # numerator
    score = start[tags[:, 0]] + emissions[ar, 0, tags[:, 0]]
    for i in range(1, L):
        score = score + (trans[tags[:, i - 1], tags[:, i]] + emissions[ar, i, tags[:, i]]) * m[:, i]
    seq_ends = mask.long().sum(dim=1) - 1
    last_tags = tags[ar, seq_ends]
    score = score + end[last_tags]
    # denominator (forward algorithm)
    z = start.unsqueeze(0) + emissions[:, 0]
    for i in range(1, L):
        nxt = torch.logsumexp(z.unsqueeze(2) + trans.unsqueeze(0) + emissions[:, i].unsqueeze(1), dim=1)
        z = torch.where(mask[:, i].bool().unsqueeze(1), nxt, z)
    logz = torch.logsumexp(z + end.unsqueeze(0), dim=1)
    return score - logz


def crf_decode(emissions: Tensor, mask: Tensor, start: Tensor, end: Tensor, trans: Tensor) -> List[List[int]]:
    """Viterbi; returns per-sample tag lists of length mask.sum() (ties -> lowest tag index, as
    torch.max/argmax-first in pytorch-crf)."""
    em = emissions.detach()
    B, L, T = em.shape
    out = []
    for b in range(B):
        n = int(mask[b].long().sum())
        score = start.detach() + em[b, 0]
        hist = []
        for i in range(1, n):
            cand = score.unsqueeze(1) + trans.detach() + em[b, i].unsqueeze(0)   # [prev, cur]
            score, idx = cand.max(dim=0)
            hist.append(idx)
        score = score + end.detach()
        best = int(score.argmax())
        tags = [best]
        for idx in reversed(hist):
            best = int(idx[best])
            tags.append(best)
        tags.reverse()
        out.append(tags)
    return out


# --------------------------------------------------------------------------------------------
# whole model: TVNetSAModel2.forward (models/bert_model.py:480-532)
# --------------------------------------------------------------------------------------------
def tvnet2_forward(p, cfg: EncoderCfg, batch: Dict[str, Tensor], *, use_prefix=True, use_probe=True,
                   vao=True, noauxloss=False, alpha=0.1, beta=0.5, past_key_values=None):
    """Eval-mode (dropout = identity) restatement of TVNetSAModel2.forward.

    batch: input_ids, attention_mask, token_type_ids, labels [B,L] int64; with use_prefix either
    images/aux_imgs/imagelabel (fusion boundary) or explicit `past_key_values` (attention boundary).
    Returns dict(loss, logits(list of lists), prob_loss, img_loss(alpha-scaled), emissions, crf_nll,
    hidden_states, prefix)."""
    ids, am, tt = batch["input_ids"], batch["attention_mask"], batch["token_type_ids"]
    B = ids.shape[0]
    img_tag_loss = torch.zeros(())
    pkv = past_key_values
    if use_prefix and pkv is None:
        pkv, img_loss, aux = visual_prompt(p, batch["images"], batch["aux_imgs"], batch.get("imagelabel"),
                                           vao=vao, n_layers=cfg.num_hidden_layers, hidden=cfg.hidden_size,
                                           n_heads=cfg.num_attention_heads)
        img_tag_loss = img_loss if noauxloss else img_loss + sum(aux)             # :489
    if pkv is not None:
        P = pkv[0][0].shape[2]
        full_mask = torch.cat([torch.ones(B, P), am.to(torch.float32)], dim=1)   # :490-492
    else:
        full_mask = am
    enc = encoder_forward(p, cfg, ids, full_mask, tt, pkv)
    hs7 = enc["hidden_states"][7] if cfg.num_hidden_layers >= 7 else enc["hidden_states"][-1]   # :503
    seq = enc["last_hidden_state"]
    emissions = F.linear(seq, p["fc.weight"], p["fc.bias"])                       # :510
    crf_p = (p["crf.start_transitions"], p["crf.end_transitions"], p["crf.transitions"])
    logits = crf_decode(emissions, am, *crf_p)                                    # :511
    out = {"emissions": emissions, "logits": logits, "hidden_states": enc["hidden_states"],
           "prefix": pkv, "position_ids": enc["position_ids"]}
    nll = None
    if "labels" in batch and batch["labels"] is not None:
        nll = -crf_log_likelihood(emissions, batch["labels"], am, *crf_p).mean()  # :521
    out["crf_nll"] = nll
    img_term = alpha * img_tag_loss
    if use_probe:
        pl, norms, labels = probe_loss(hs7, p["oneWordpsdProbe.oneWordpsdProbe.proj"])   # :508
        out.update(prob_loss=pl, norms=norms, pseudo_labels=labels)
        out["loss"] = combine_loss(nll, pl, beta, 30) + img_term                  # :523-525
    else:
        out["loss"] = nll + img_term                                              # :530
    out["img_loss"] = img_term
    return out


# --------------------------------------------------------------------------------------------
# span variant: TVNetSAModel.forward / extraction / classification (models/bert_model.py:246-376)
# --------------------------------------------------------------------------------------------
def span_representation(span_starts: Tensor, span_ends: Tensor, x: Tensor, input_mask: Tensor):
    """get_span_representation (models/bert_model.py:147-172): spans index the COMPACTED token stream (all
    tokens with mask 1, sentence after sentence), clamped to its last element; JR = widest span."""
    mask = input_mask.to(span_starts.dtype)
    input_len = mask.sum(-1)
    word_offset = torch.cumsum(input_len, 0) - input_len
    s = (span_starts + word_offset.unsqueeze(1)).reshape(-1)
    e = (span_ends + word_offset.unsqueeze(1)).reshape(-1)
    width = e - s + 1
    JR = int(width.max())
    B, L, H = x.shape
    flat = x.reshape(B * L, H)[mask.reshape(-1).nonzero().squeeze(-1), :]        # flatten_emb_by_sentence :140-145
    total = flat.shape[0]
    idx = torch.arange(JR).unsqueeze(0) + s.unsqueeze(1)
    idx = torch.minimum(idx, torch.full_like(idx, total - 1))                   # :165
    span_emb = flat[idx, :]                                                      # [N*M, JR, H]
    span_mask = torch.arange(JR).unsqueeze(0) < width.unsqueeze(-1)
    return span_emb, span_mask


def self_att_representation(x: Tensor, score: Tensor, mask: Tensor) -> Tensor:
    """get_self_att_representation (:174-181): softmax(score + (1-mask) * -10000) weighted sum."""
    m = (1.0 - mask.to(score.dtype)) * -10000.0
    prob = torch.softmax(score + m, dim=-1).unsqueeze(-1)
    return (prob * x).sum(1)


def distant_cross_entropy(logits: Tensor, positions: Tensor) -> Tensor:
    """distant_cross_entropy without mask (:183-192)."""
    lp = torch.log_softmax(logits, dim=-1)
    pos = positions.to(lp.dtype)
    return -torch.mean((pos * lp).sum(-1) / pos.sum(-1))


def tvnet_forward(p, cfg: EncoderCfg, batch: Dict[str, Tensor], *, use_prefix=True, use_probe=True, beta=0.5,
                  num_epochs=30):
    """Eval-mode restatement of TVNetSAModel.forward (models/bert_model.py:246-321; no GCN, no Cutoff).

    batch: input_ids, attention_mask, token_type_ids [B,L]; start_positions / end_positions [B,L] (multi-hot);
    span_starts / span_ends [B,M]; polarity_labels / label_masks [B,M]; images / aux_imgs with use_prefix.
    Returns dict(loss, tot_loss, prob_loss, start_logits, end_logits, ac_logits, logits)."""
    ids, am, tt = batch["input_ids"], batch["attention_mask"], batch["token_type_ids"]
    B = ids.shape[0]
    pkv = None
    if use_prefix:
        pkv, _, _ = visual_prompt(p, batch["images"], batch["aux_imgs"], None, vao=False,
                                  n_layers=cfg.num_hidden_layers, hidden=cfg.hidden_size,
                                  n_heads=cfg.num_attention_heads)               # :379-414 (no ANP heads)
        P = pkv[0][0].shape[2]
        full_mask = torch.cat([torch.ones(B, P), am.to(torch.float32)], dim=1)   # :257-259
    else:
        full_mask = am
    enc = encoder_forward(p, cfg, ids, full_mask, tt, pkv)
    seq = enc["last_hidden_state"]                                               # dropout = identity in eval
    ae = F.linear(seq, p["binary_affine.weight"], p["binary_affine.bias"])      # :351
    start_logits, end_logits = ae[..., 0], ae[..., 1]
    span_emb, span_mask = span_representation(batch["span_starts"], batch["span_ends"], seq, am)   # :364
    score = F.linear(span_emb, p["unary_affine.weight"], p["unary_affine.bias"]).squeeze(-1)       # :367-368
    pooled = self_att_representation(span_emb, score, span_mask)                 # :369
    pooled = torch.tanh(F.linear(pooled, p["dense.weight"], p["dense.bias"]))    # :371-372
    ac_logits = F.linear(pooled, p["classifier.weight"], p["classifier.bias"])   # :374
    flat_labels = batch["polarity_labels"].reshape(-1)
    flat_masks = batch["label_masks"].reshape(-1).to(ac_logits.dtype)
    start_loss = distant_cross_entropy(start_logits, batch["start_positions"])   # :298-300
    end_loss = distant_cross_entropy(end_logits, batch["end_positions"])
    ae_loss = (start_loss + end_loss) / 2
    ac_loss = F.cross_entropy(ac_logits, flat_labels)                            # mean over all spans (:302)
    ac_loss = torch.sum(flat_masks * ac_loss) / flat_masks.sum()                 # :303 (a scalar times mask: no-op)
    tot = ae_loss + ac_loss
    out = {"start_logits": start_logits, "end_logits": end_logits, "ac_logits": ac_logits,
           "logits": ac_logits.view(B, -1, ac_logits.shape[-1]), "tot_loss": tot,
           "hidden_states": enc["hidden_states"]}
    if use_probe:
        pl, norms, labels = probe_loss(enc["hidden_states"][7], p["oneWordpsdProbe.oneWordpsdProbe.proj"])   # :356
        out.update(prob_loss=pl, loss=combine_loss(tot, pl, beta, num_epochs))   # :312
    else:
        out["loss"] = tot
    return out
