"""Drop-in nn.Module surface of the hot path.

Same class names, constructor / forward signatures, attribute and state_dict names (and order) as the
reference (`models/modeling_roberta.py`, `models/modeling_bert.py`, `models/bert_model.py`, `probes/`),
but every forward/backward runs through the sm_100a kernels of this package (mtvaf_b200.engine).
The torch.nn sub-modules below are PARAMETER CONTAINERS only: their own forward() is never used.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from types import SimpleNamespace
from typing import List, Optional

import torch
from torch import nn

from . import lib as L
from . import ops
from .engine import Engine, HotPathConfig, BF16, F32


def _resolve_dtype(x) -> torch.dtype:
    if x is None:
        x = os.environ.get("MTVAF_COMPUTE", "bf16")
    if isinstance(x, torch.dtype):
        return x
    x = str(x).lower()
    if x in ("bf16", "bfloat16"):
        return BF16
    if x in ("fp32", "float32", "f32"):
        return F32
    raise ValueError("compute dtype must be bf16 or fp32, got %r" % (x,))


class ModelOutput(OrderedDict):
    """Indexable by key, attribute or position like transformers' ModelOutput (None entries skipped
    for positional access, as the reference's callers expect: models/bert_model.py:324-349,496-505)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __getitem__(self, k):
        if isinstance(k, (int, slice)):
            vals = [v for v in self.values() if v is not None]
            return tuple(vals)[k] if isinstance(k, slice) else vals[k]
        return OrderedDict.__getitem__(self, k)

    def to_tuple(self):
        return tuple(v for v in self.values() if v is not None)


def TokenClassifierOutput(loss=None, logits=None, hidden_states=None, attentions=None):
    return ModelOutput(loss=loss, logits=logits, hidden_states=hidden_states, attentions=attentions)


# =================================================================================================
# encoder containers (names/order match models/modeling_roberta.py:76-100,174-176,291-292,358,374-375)
# =================================================================================================
class _Embeddings(nn.Module):
    def __init__(self, config, roberta: bool):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size,
                                                padding_idx=config.pad_token_id if roberta else None)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.register_buffer("token_type_ids", torch.zeros((1, config.max_position_embeddings), dtype=torch.long),
                             persistent=False)


class _SelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        H = config.hidden_size
        self.query, self.key, self.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class _SelfOutput(nn.Module):
    def __init__(self, config, in_features):
        super().__init__()
        self.dense = nn.Linear(in_features, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class _Attention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = _SelfAttention(config)
        self.output = _SelfOutput(config, config.hidden_size)


class _Intermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class _Layer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = _Attention(config)
        self.intermediate = _Intermediate(config)
        self.output = _SelfOutput(config, config.intermediate_size)


class _Encoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([_Layer(config) for _ in range(config.num_hidden_layers)])


class _Pooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()


def _init_like_reference(module: nn.Module, std: float):
    """_init_weights models/modeling_roberta.py:695-709."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)


# =================================================================================================
# autograd bridges
# =================================================================================================
_ANCHORS = {}


def _grad_anchor(like: torch.Tensor) -> torch.Tensor:
    """Autograd anchor of the hand-written Functions: a HOST leaf that requires grad iff `like` (a representative
    parameter) does.  The Functions write parameter gradients straight into the flat buffer, so they need some
    differentiable input to be scheduled at all; a CUDA parameter in that role drags its cached AccumulateGrad
    node -- and the stream it was first used on -- into every later backward, which breaks CUDA-graph capture
    on another stream ("dependency created on uncaptured work").  A host tensor carries no stream."""
    key = bool(like.requires_grad)
    if key not in _ANCHORS:
        _ANCHORS[key] = torch.zeros(1, requires_grad=key)
    return _ANCHORS[key]


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: Engine, ids, tts, mask, training, want_probs, embeds, anchor, kv):
        need_grad = any(ctx.needs_input_grad)       # grad mode is off inside Function.forward
        ctx.set_materialize_grads(False)            # unused hidden states arrive as None, not as zero tensors
        hs, saved, attns = engine.encoder_fwd(ids, tts, mask, kv, training, need_grad, want_probs, embeds)
        ctx.engine, ctx.saved = engine, saved
        ctx.kv_grad = kv is not None and kv.requires_grad
        ctx.kv_dtype = None if kv is None else kv.dtype
        # prefix produced by this model's own fusion stack in the same forward (single consumer): its fp32 gradient is
        # handed to _FusionFn.backward directly instead of through a bf16 round trip (600 MB each way at B=512)
        ctx.kv_ptr = None
        if kv is not None and getattr(engine, "_kv_internal_ptr", None) == kv.data_ptr():
            ctx.kv_ptr = kv.data_ptr()
            engine._kv_internal_ptr = None
        ctx.emb_grad = embeds is not None and embeds.requires_grad
        B, Lq = ids.shape
        outs = tuple(h.view(B, Lq, -1) for h in hs)
        if want_probs:
            ctx.mark_non_differentiable(*attns)
            return outs + tuple(attns)
        return outs

    @staticmethod
    def backward(ctx, *grads):
        eng = ctx.engine
        n = eng.cfg.n_layers
        eng.flat.attach_grads()
        dkv, demb = eng.encoder_bwd(ctx.saved, list(grads[:n + 1]), ctx.kv_grad, ctx.emb_grad)
        ctx.saved = None
        if dkv is not None and ctx.kv_dtype == BF16:
            if ctx.kv_ptr is not None:
                eng._dkv32[ctx.kv_ptr] = dkv
                if getattr(eng, "_dummy_bf16", None) is None or eng._dummy_bf16.device != dkv.device:
                    eng._dummy_bf16 = torch.zeros(1, dtype=BF16, device=dkv.device)
                dkv = eng._dummy_bf16.expand(dkv.shape)      # placeholder of the right shape / dtype for autograd
            else:
                dkv = ops.cast_bf16(dkv)
        # (`embeds` entered as [T, H]; its gradient leaves in the same shape)
        return (None, None, None, None, None, None, demb, None, dkv)


class _FusionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: Engine, feats, imagelabel, vao, training, n_aux_heads, anchor):
        need_grad = any(ctx.needs_input_grad)
        kv, img_losses, saved = engine.fusion_fwd(feats, imagelabel, vao, training, need_grad, n_aux_heads)
        ctx.engine, ctx.saved = engine, saved
        ctx.kv_ptr = kv.data_ptr()
        if img_losses is None:
            img_losses = torch.zeros(feats.shape[0], dtype=F32, device=feats.device)
        return kv, img_losses

    @staticmethod
    def backward(ctx, dkv, dimg):
        eng = ctx.engine
        eng.flat.attach_grads()
        d32 = eng._dkv32.pop(ctx.kv_ptr, None)               # fp32 gradient handed over by _EncoderFn.backward
        if d32 is not None:
            dkv = d32
        if dkv is None:
            dkv = torch.zeros_like(ctx.saved["gates"]).new_zeros(
                (eng.cfg.n_layers, 2, ctx.saved["B"], 4 * ctx.saved["n_img"] * eng.cfg.H))
        if dkv.dtype != F32:
            dkv = ops.cast_f32(dkv.contiguous())
        eng.fusion_bwd(ctx.saved, dkv.contiguous(), None if dimg is None else dimg.contiguous())
        ctx.saved = None
        return (None,) * 7


class _HeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: Engine, model, hs_probe, hs_last, mask, labels, img_losses, training):
        a = model.args
        B, Lq = mask.shape
        H = engine.cfg.H
        hs = {engine.cfg.n_layers: hs_last.reshape(B * Lq, H), 7: None}
        probe_layer = min(7, engine.cfg.n_layers)
        hs[probe_layer] = hs_probe.reshape(B * Lq, H) if hs_probe is not None else None
        need_grad = any(ctx.needs_input_grad) and labels is not None
        ctx.set_materialize_grads(False)             # d(prob_loss) arrives as None unless someone backprops through it
        use_probe = bool(getattr(a, "use_probe", False))
        counted = None
        if img_losses is not None:
            counted = img_losses[:1] if getattr(a, "noauxloss", False) else img_losses
            counted = counted.contiguous()
        out, saved = engine.heads_fwd(hs, B, Lq, mask, labels, use_probe, float(getattr(a, "beta", 0.5)),
                                      float(getattr(a, "alpha", 0.0)), counted, training, need_grad,
                                      probe_layer=probe_layer)
        ctx.engine, ctx.saved = engine, saved
        ctx.n_img = 0 if img_losses is None else img_losses.numel()
        ctx.n_counted = 0 if counted is None else counted.numel()
        ctx.alpha = float(getattr(a, "alpha", 0.0))
        ctx.shape = hs_last.shape
        ctx.has_probe_in = hs_probe is not None
        model._last_heads = out
        loss = out["loss"] if out["loss"] is not None else torch.zeros(1, dtype=F32, device=mask.device)
        prob = out["prob_loss"] if out["prob_loss"] is not None else torch.zeros(1, dtype=F32, device=mask.device)
        # prob_loss stays attached to autograd like the reference's: its gradient joins the probe
        # branch of the heads backward
        return loss.view(()), prob.view(())

    @staticmethod
    def backward(ctx, dloss, dprob):
        eng = ctx.engine
        if ctx.saved is None:
            raise L.MtvafError("backward through the heads needs labels (no loss was computed)")
        eng.flat.attach_grads()
        dev = ctx.saved["seq_d"].device
        dl = dloss.reshape(1).to(F32).contiguous() if dloss is not None else torch.zeros(1, dtype=F32, device=dev)
        dp = None if dprob is None else dprob.reshape(1).to(F32).contiguous()
        grads = eng.heads_bwd(ctx.saved, dl, dp)
        n = eng.cfg.n_layers
        d_last = grads[n].view(ctx.shape)
        d_probe = None
        pl = ctx.saved["probe_layer"]
        if ctx.has_probe_in and pl in grads and pl != n:
            d_probe = grads[pl].view(ctx.shape)
        elif ctx.has_probe_in and pl == n:
            raise L.MtvafError("probe layer must differ from the last layer (needs >= 8 encoder layers)")
        d_img = None
        if ctx.n_img:
            d_img = torch.zeros(ctx.n_img, dtype=F32, device=dl.device)
            d_img[:ctx.n_counted] = dl * ctx.alpha
        ctx.saved = None
        return (None, None, d_probe, d_last, None, None, d_img, None)


# =================================================================================================
# RobertaModel / BertModel
# =================================================================================================
class _EncoderModelBase(nn.Module):
    KIND = "roberta"

    def __init__(self, config, add_pooling_layer=True, compute_dtype=None):
        super().__init__()
        self.config = config
        roberta = self.KIND == "roberta"
        self.embeddings = _Embeddings(config, roberta)
        self.encoder = _Encoder(config)
        self.pooler = _Pooler(config) if add_pooling_layer else None
        _init_like_reference(self, getattr(config, "initializer_range", 0.02))
        self._engine: Optional[Engine] = None
        self._engine_owner = None          # set by a wrapping model so one Engine covers all its params
        self._engine_prefix = ""
        self._compute_dtype = _resolve_dtype(compute_dtype)

    # -- construction helpers -----------------------------------------------------------------
    @classmethod
    def from_config(cls, config, **kw):
        return cls(config, **kw)

    @classmethod
    def from_pretrained(cls, name_or_path, *a, **kw):
        """Loads weights through transformers' own loader when they are available locally.  There is no
        network in the box: for benchmarks / tests use `from_config` (random init) instead."""
        import transformers
        hf_cls = transformers.RobertaModel if cls.KIND == "roberta" else transformers.BertModel
        try:
            hf = hf_cls.from_pretrained(name_or_path, *a, **kw)
        except Exception as exc:   # offline
            raise L.MtvafError("cannot load %r offline (%s); build the encoder with %s.from_config(config)"
                               % (name_or_path, type(exc).__name__, cls.__name__)) from exc
        model = cls(hf.config)
        missing = model.load_state_dict(hf.state_dict(), strict=False)
        bad = [k for k in missing.missing_keys if "position_ids" not in k and "token_type_ids" not in k]
        if bad:
            raise L.MtvafError("state_dict mismatch: %s" % bad)
        return model

    # -- engine ---------------------------------------------------------------------------------
    def hot_config(self) -> HotPathConfig:
        c = self.config
        return HotPathConfig(self.KIND, c.hidden_size, c.num_attention_heads, c.intermediate_size,
                             c.num_hidden_layers, c.layer_norm_eps, c.pad_token_id,
                             getattr(c, "hidden_dropout_prob", 0.1), getattr(c, "attention_probs_dropout_prob", 0.1))

    def set_compute_dtype(self, dtype):
        self._compute_dtype = _resolve_dtype(dtype)
        if self._engine is not None:
            self._engine.compute_dtype = self._compute_dtype

    def engine(self) -> Engine:
        if self._engine_owner is not None:
            return self._engine_owner()
        if self._engine is None:
            self._engine = Engine(self, self.hot_config(), "", self._compute_dtype)
        return self._engine

    def _anchor(self) -> torch.Tensor:
        return _grad_anchor(self.embeddings.LayerNorm.weight)

    # -- reference API ----------------------------------------------------------------------------
    def _pack_prefix(self, past_key_values, B, eng: Engine):
        if past_key_values is None:
            return None
        if torch.is_tensor(past_key_values):
            return past_key_values           # already packed [n_layers,2,B,P*H] by the fusion stack
        # list of n_layers (K,V) [B,heads,P,d] (attention boundary): stacking/casting is memory plumbing
        ks = torch.stack([torch.stack([k.reshape(B, -1), v.reshape(B, -1)]) for k, v in past_key_values])
        if ks.dtype != eng.compute_dtype:
            ks = ks.to(eng.compute_dtype)
        return ks.contiguous()

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None, past_key_values=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        """models/modeling_roberta.py:850-978.  `attention_mask` is [B, P+L] when a prefix is given
        (prefix columns first, all ones: models/bert_model.py:490-492)."""
        if input_ids is None:
            raise L.MtvafError("mtvaf_b200 encoder needs input_ids (inputs_embeds goes through get_bert_output)")
        if position_ids is not None or head_mask is not None or encoder_hidden_states is not None:
            raise L.MtvafError("position_ids / head_mask / cross-attention inputs are not on the MTVAF path")
        if not input_ids.is_cuda:
            raise L.MtvafError("mtvaf_b200 runs on CUDA devices only (no CPU fallback); got %s" % input_ids.device)
        eng = self.engine()
        eng.compute_dtype = self._compute_dtype if self._engine_owner is None else eng.compute_dtype
        eng.prepare()
        eng.new_step()                       # fresh dropout masks per call (no-op inside a wrapping model's forward)
        B, Lq = input_ids.shape
        kv = self._pack_prefix(past_key_values, B, eng)
        P = 0 if kv is None else kv.shape[3] // self.config.hidden_size
        if attention_mask is None:
            text_mask = torch.ones((B, Lq), dtype=torch.long, device=input_ids.device)
        else:
            if attention_mask.shape[1] == P + Lq:
                text_mask = attention_mask[:, P:]
            elif attention_mask.shape[1] == Lq:
                text_mask = attention_mask
            else:
                raise L.MtvafError("attention_mask has %d columns, expected %d (prefix + text)"
                                   % (attention_mask.shape[1], P + Lq))
            text_mask = text_mask.to(torch.long).contiguous()
        if token_type_ids is None:
            token_type_ids = torch.zeros((B, Lq), dtype=torch.long, device=input_ids.device)
        want_probs = bool(output_attentions) and bool(getattr(self.config, "materialize_attentions", False))
        outs = _EncoderFn.apply(eng, input_ids.contiguous(), token_type_ids.contiguous(), text_mask, self.training,
                                want_probs, None, self._anchor(), kv)
        n = self.config.num_hidden_layers
        hs = outs[:n + 1]
        attns = tuple(outs[n + 1:]) if want_probs else None
        return self._wrap(hs, attns, output_hidden_states)

    def _wrap(self, hs, attns, output_hidden_states=True):
        last = hs[-1]
        # pooler (models/modeling_roberta.py:675-681) is dead on the MTVAF path: lazy
        pooled = _LazyPooler(self, last) if self.pooler is not None else None
        return ModelOutput(last_hidden_state=last, pooler_output=pooled, hidden_states=tuple(hs), attentions=attns)

    def pooled(self, last_hidden_state: torch.Tensor) -> torch.Tensor:
        """tanh(W h[:,0] + b) -- inference-only helper (no gradient: the path never trains the pooler)."""
        eng = self.engine()
        with torch.no_grad():
            first = last_hidden_state[:, 0].contiguous()
            name = (eng.enc_prefix + "pooler.dense.weight")
            return ops.linear_fwd(first, eng.cw(name), eng.flat.w(eng.enc_prefix + "pooler.dense.bias"),
                                  mode=L.EPI_TANH)

    def get_embedding_output(self, input_ids, token_type_ids=None, position_ids=None):
        """models/modeling_roberta.py:980-988 (used by Cutoff, modules/augument.py:61).  Differentiable like the
        reference's: the embedding tables and the embedding LayerNorm receive gradients through it."""
        if position_ids is not None:
            raise L.MtvafError("explicit position_ids are not on the MTVAF path")
        eng = self.engine()
        eng.prepare()
        eng.new_step()
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        B, Lq = input_ids.shape
        x = _EmbedFn.apply(eng, input_ids.contiguous(), token_type_ids.contiguous(), self.training, self._anchor())
        return x.view(B, Lq, -1)

    def get_bert_output(self, embedding_output, attention_mask=None, past_key_values=None):
        """models/modeling_roberta.py:990-1020: returns (sequence_output, pooled_output, attentions).  Gradients flow
        back into `embedding_output` (Cutoff trains through it, modules/augument.py:75)."""
        eng = self.engine()
        eng.prepare()
        eng.new_step()
        B, Lq, H = embedding_output.shape
        kv = self._pack_prefix(past_key_values, B, eng)
        P = 0 if kv is None else kv.shape[3] // H
        if attention_mask is None:
            text_mask = torch.ones((B, Lq), dtype=torch.long, device=embedding_output.device)
        elif attention_mask.shape[1] == P + Lq:
            # the kernels attend to every prefix row (models/bert_model.py:490-492 builds those columns as ones);
            # a caller that zeroes prefix columns must hear about it instead of getting silently different math
            if P and not bool((attention_mask[:, :P] != 0).all()):
                raise L.MtvafError("get_bert_output: masked-out prefix columns are not supported by the fused "
                                   "prefix attention (all %d prefix columns must be 1)" % P)
            text_mask = attention_mask[:, P:]
        elif attention_mask.shape[1] == Lq:
            text_mask = attention_mask
        else:
            raise L.MtvafError("attention_mask has %d columns, expected %d (prefix + text)"
                               % (attention_mask.shape[1], P + Lq))
        ids = torch.zeros((B, Lq), dtype=torch.long, device=embedding_output.device)
        emb = embedding_output.reshape(B * Lq, H)
        if emb.dtype != eng.compute_dtype:
            emb = emb.to(eng.compute_dtype)          # differentiable cast (plumbing)
        outs = _EncoderFn.apply(eng, ids, ids, text_mask.to(torch.long).contiguous(), self.training, False,
                                emb.contiguous(), self._anchor(), kv)
        last = outs[self.config.num_hidden_layers]
        return (last, _LazyPooler(self, last), None)


class _EmbedFn(torch.autograd.Function):
    """Embedding gather + LayerNorm (+dropout) as its own autograd node (get_embedding_output)."""

    @staticmethod
    def forward(ctx, engine: Engine, ids, tts, training, anchor):
        c, f, e = engine.cfg, engine.flat, engine.enc_prefix
        p_h = c.hidden_dropout if training else 0.0
        seed = engine.seed(1)
        x, pids, mean, rstd = ops.embed_ln_fwd(ids, tts, f.w(e + "embeddings.word_embeddings.weight"),
                                               f.w(e + "embeddings.position_embeddings.weight"),
                                               f.w(e + "embeddings.token_type_embeddings.weight"),
                                               f.w(e + "embeddings.LayerNorm.weight"),
                                               f.w(e + "embeddings.LayerNorm.bias"), c.eps,
                                               0 if c.kind == "roberta" else 1, c.pad_id, engine.compute_dtype, p_h, seed)
        ctx.engine, ctx.saved = engine, (ids, tts, pids, mean, rstd, p_h, seed)
        return x

    @staticmethod
    def backward(ctx, dx):
        eng = ctx.engine
        c, f, e = eng.cfg, eng.flat, eng.enc_prefix
        ids, tts, pids, mean, rstd, p_h, seed = ctx.saved
        f.attach_grads()
        dx = dx.reshape(ids.numel(), c.H).contiguous()
        if dx.dtype != eng.compute_dtype:
            dx = ops.cast_bf16(dx) if eng.compute_dtype == BF16 else ops.cast_f32(dx)
        ops.embed_ln_bwd(dx, ids, tts, pids, f.w(e + "embeddings.word_embeddings.weight"),
                         f.w(e + "embeddings.position_embeddings.weight"),
                         f.w(e + "embeddings.token_type_embeddings.weight"),
                         f.w(e + "embeddings.LayerNorm.weight"), mean, rstd, 0 if c.kind == "roberta" else 1, c.pad_id,
                         f.g(e + "embeddings.word_embeddings.weight"), f.g(e + "embeddings.position_embeddings.weight"),
                         f.g(e + "embeddings.token_type_embeddings.weight"), f.g(e + "embeddings.LayerNorm.weight"),
                         f.g(e + "embeddings.LayerNorm.bias"), p_h, seed)
        ctx.saved = None
        return (None,) * 5


class _LazyPooler:
    """Materialises the (dead) pooler output only if someone reads it."""

    def __init__(self, model, last):
        self._m, self._last, self._v = model, last, None

    def tensor(self) -> torch.Tensor:
        if self._v is None:
            self._v = self._m.pooled(self._last)
        return self._v

    def __getattr__(self, k):
        return getattr(self.tensor(), k)

    def __getitem__(self, i):
        return self.tensor()[i]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        """torch.* functions see the materialised tensor (the reference returns a plain Tensor here)."""
        un = lambda a: a.tensor() if isinstance(a, _LazyPooler) else a
        args = tuple(un(a) for a in args)
        kwargs = {k: un(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)


class RobertaModel(_EncoderModelBase):
    KIND = "roberta"


class BertModel(_EncoderModelBase):
    KIND = "bert"


# =================================================================================================
# probes (probes/probe.py, probes/constructLabel.py, probes/probe_trainModel.py, probes/loss.py)
# =================================================================================================
class OneWordPSDProbe(nn.Module):
    """probes/probe.py:50-79."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.probe_rank = args["probe"]["maximum_rank"]
        self.model_dim = args["model"]["hidden_dim"]
        self.proj = nn.Parameter(torch.zeros(self.model_dim, self.probe_rank))
        nn.init.uniform_(self.proj, -0.05, 0.05)

    def forward(self, batch):
        """Squared L2 norm of batch @ proj per token (inference helper; the training path goes through
        the fused heads of TVNetSAModel2)."""
        B, Lq, H = batch.shape
        with torch.no_grad():
            x = batch.reshape(B * Lq, H).contiguous()
            proj = self.proj.detach()
            if x.dtype == BF16:
                proj = ops.cast_bf16(proj.contiguous())
            norms = torch.zeros(B * Lq, dtype=F32, device=batch.device)
            ops.gemm(x, proj, b_mn=True, M=B * Lq, N=self.probe_rank, K=H, mode=L.EPI_SQNORM, rowvec=norms)
        return norms.view(B, Lq)


class TwoWordPSDProbe(nn.Module):
    """probes/probe.py:9-46 (defined but never called by the reference; kept for the microbenchmarks)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.probe_rank = args["probe"]["maximum_rank"]
        self.model_dim = args["model"]["hidden_dim"]
        self.proj = nn.Parameter(torch.zeros(self.model_dim, self.probe_rank))
        nn.init.uniform_(self.proj, -0.05, 0.05)

    def forward(self, batch):
        B, Lq, H = batch.shape
        with torch.no_grad():
            x = batch.reshape(B * Lq, H).contiguous()
            proj = self.proj.detach()
            if x.dtype == BF16:
                proj = ops.cast_bf16(proj.contiguous())
            T = ops.gemm(x, proj, b_mn=True, M=B * Lq, N=self.probe_rank, K=H, out_dtype=F32)
            return ops.pairwise_sqdist(T, B, Lq, self.probe_rank)


class ConstructLabelGaget(nn.Module):
    """probes/constructLabel.py:6-29, on the device (bit-exact), no host loop / sync storm."""

    def __init__(self, args=None):
        super().__init__()

    def forward(self, norms):
        return ops.probe_labels(norms.detach().to(F32).contiguous())


class probe(nn.Module):
    """probes/probe_trainModel.py:9-26."""

    def __init__(self, args):
        super().__init__()
        self.oneWordpsdProbe = OneWordPSDProbe(args={"probe": {"maximum_rank": args["probe"]["maximum_rank"]},
                                                     "model": {"hidden_dim": args["model"]["hidden_dim"]}})
        self.constructLabel = ConstructLabelGaget(args=None)

    def forward(self, batch):
        norms = self.oneWordpsdProbe(batch)
        labels = self.constructLabel(norms)
        loss, _ = ops.mse(norms.reshape(-1).contiguous(), labels.reshape(-1), False)
        return loss.view(())


class CombineLoss(nn.Module):
    """probes/loss.py:4-18 with the `.item()` predicate evaluated on the device."""

    def __init__(self, para):
        super().__init__()
        self.superParameter = torch.tensor(para)

    def forward(self, loss, probe_loss, epoch):
        # scalar glue kept differentiable for stand-alone use; the model's own forward uses the fused
        # mtvaf_combine_loss kernel instead
        coef = float(self.superParameter) * 2.0 ** (-int(epoch))
        return loss + torch.where(probe_loss > 0.1, probe_loss * coef, torch.zeros_like(probe_loss))


# =================================================================================================
# CRF (pytorch-crf API; call sites models/bert_model.py:464,511,521)
# =================================================================================================
class CRF(nn.Module):
    def __init__(self, num_tags: int, batch_first: bool = False):
        super().__init__()
        self.num_tags, self.batch_first = num_tags, batch_first
        self.start_transitions = nn.Parameter(torch.empty(num_tags))
        self.end_transitions = nn.Parameter(torch.empty(num_tags))
        self.transitions = nn.Parameter(torch.empty(num_tags, num_tags))
        for p in (self.start_transitions, self.end_transitions, self.transitions):
            nn.init.uniform_(p, -0.1, 0.1)

    def _prep(self, emissions, mask):
        if not self.batch_first:
            emissions = emissions.transpose(0, 1)
            mask = None if mask is None else mask.transpose(0, 1)
        B, Lq, _ = emissions.shape
        if mask is None:
            mask = torch.ones((B, Lq), dtype=torch.long, device=emissions.device)
        return emissions.to(F32).contiguous(), mask.to(torch.long).contiguous()

    def decode(self, emissions, mask=None) -> List[List[int]]:
        em, m = self._prep(emissions.detach(), mask)
        best, lens = ops.crf_decode(em, m, self.start_transitions.detach(), self.end_transitions.detach(),
                                    self.transitions.detach())
        best, lens = best.cpu(), lens.cpu()
        return [best[b, :int(lens[b])].tolist() for b in range(best.shape[0])]

    def forward(self, emissions, tags, mask=None, reduction="sum"):
        em, m = self._prep(emissions, mask)
        tg = tags if self.batch_first else tags.transpose(0, 1)
        return _CrfFn.apply(em, tg.to(torch.long).contiguous(), m, self.start_transitions, self.end_transitions,
                            self.transitions, reduction)


class _CrfFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, em, tags, mask, start, end, trans, reduction):
        B = em.shape[0]
        if reduction not in ("sum", "mean"):
            raise L.MtvafError("CRF reduction %r not supported on the device path" % reduction)
        scale = 1.0 / B if reduction == "mean" else 1.0
        nll, d_em, d_s, d_e, d_t = ops.crf_nll(em, tags, mask, start.detach(), end.detach(), trans.detach(), True,
                                               scale)
        ctx.save_for_backward(d_em, d_s, d_e, d_t)
        return (-nll * scale).view(())

    @staticmethod
    def backward(ctx, g):
        d_em, d_s, d_e, d_t = ctx.saved_tensors
        return -g * d_em, None, None, -g * d_s, -g * d_e, -g * d_t, None


# =================================================================================================
# TVNetSAModel2 (models/bert_model.py:416-588)
# =================================================================================================
class ImageModel(nn.Module):
    """Frozen ResNet pyramid (models/bert_model.py:63-111).  OUT OF SCOPE for kernels (SURVEY.md section 2
    row 5): kept as the reference's torchvision wrapper so real images still work; the hot path starts
    at its outputs.  `FeatureStub` below is used when pyramid features are fed directly."""

    def __init__(self, use_152=False, use_101=False, use_34=False, use_18=False, resnet_root=None):
        super().__init__()
        import torchvision.models as tvm
        name = "resnet152" if use_152 else "resnet101" if use_101 else "resnet34" if use_34 else \
            "resnet18" if use_18 else "resnet50"
        self.resnet = getattr(tvm, name)(weights=None)
        # the reference loads resnet_root/<name>.pth unconditionally (models/bert_model.py:84-85).  A mistyped path
        # must not leave a randomly initialised frozen front-end behind: raise unless the caller opts out explicitly
        # (resnet_root=None: offline / random-init runs, e.g. the synthetic benchmarks)
        if resnet_root is not None:
            path = os.path.join(resnet_root, name + ".pth")
            if not os.path.exists(path):
                raise FileNotFoundError("ImageModel: %s not found (pass resnet_root=None for a random-init "
                                        "front-end)" % path)
            self.resnet.load_state_dict(torch.load(path))

    def forward(self, x, aux_imgs=None):
        main = self.get_resnet_prompt(x)
        if aux_imgs is None:
            return main, None
        aux = aux_imgs.permute([1, 0, 2, 3, 4])
        return main, [self.get_resnet_prompt(aux[i]) for i in range(len(aux))]

    def get_resnet_prompt(self, x):
        out = []
        for name, layer in self.resnet.named_children():
            if name in ("fc", "avgpool"):
                continue
            x = layer(x)
            if "layer" in name:
                k = x.size(2) // 2
                out.append(nn.functional.avg_pool2d(x, kernel_size=(k, k), stride=k))
        return out


class FeatureStub(nn.Module):
    """Takes packed pyramid features images [B,3840,2,2], aux_imgs [B,n_aux,3840,2,2] (the fusion
    boundary of SURVEY.md 8(d); fp32, or bf16 = the cached wire format of mtvaf_b200.features) and returns them in
    ImageModel's list-of-4 layout.  The models recognise `packed_features` and hand the tensors to
    `mtvaf_pack_features` directly instead of going through split / cat / stack."""
    WIDTHS = (256, 512, 1024, 2048)
    packed_features = True

    def forward(self, x, aux_imgs=None):
        main = list(torch.split(x, self.WIDTHS, dim=1))
        if aux_imgs is None:
            return main, None
        aux = aux_imgs.permute(1, 0, 2, 3, 4)
        return main, [list(torch.split(aux[i], self.WIDTHS, dim=1)) for i in range(aux.shape[0])]


class TVNetSAModel2(nn.Module):
    def __init__(self, label_list, tokenizer, args, type_num=None, use_weight=False, config=None,
                 image_model: Optional[nn.Module] = None):
        """Same signature as models/bert_model.py:417; two optional extras for offline use:
        `config` builds a random-init encoder instead of from_pretrained(args.bert_name); `image_model`
        replaces the torchvision ResNet (e.g. FeatureStub() to feed pyramid features directly)."""
        super().__init__()
        self.args = args
        self.type_num = type_num
        self.tokenizer = tokenizer
        self.prefix_dim = args.prefix_dim
        self.prefix_len = args.prefix_len
        enc_cls = RobertaModel if "roberta" in args.bert_name else BertModel
        self.bert = enc_cls.from_config(config) if config is not None else enc_cls.from_pretrained(args.bert_name)
        H = self.bert.config.hidden_size
        nl = self.bert.config.num_hidden_layers
        self.num_labels = len(label_list) + 1
        if args.use_prefix:
            if image_model is not None:
                self.image_model = image_model
            else:
                self.image_model = ImageModel(use_152=getattr(args, "use_152", False),
                                              use_101=getattr(args, "use_101", False),
                                              use_34=getattr(args, "use_34", False),
                                              use_18=getattr(args, "use_18", False), resnet_root=args.resnet_root)
            small = getattr(args, "use_34", False) or getattr(args, "use_18", False)
            self.encoder_conv = nn.Sequential(nn.Linear(960 if small else 3840, 800), nn.Tanh(),
                                              nn.Linear(800, 4 * 2 * H))
            self.projectors = nn.ModuleList([nn.Linear(4 * H * 2, 4) for _ in range(nl)])
            self.img_dropout = nn.Dropout(0.2)
            self.img_classifier = nn.Linear(4 * 2 * H, 2089)
            self.aux_img_classifier = nn.ModuleList([nn.Linear(4 * 2 * H, 2089) for _ in range(3)])
        self.crf = CRF(self.num_labels, batch_first=True)
        self.fc = nn.Linear(H, self.num_labels)
        self.dropout = nn.Dropout(0.1)
        if args.use_probe:
            self.oneWordpsdProbe = probe(args={"probe": {"maximum_rank": H // 2}, "model": {"hidden_dim": H}})
            # models/bert_model.py:474-475 loads ./models/psdProbe_base_savel7.pt unconditionally.  `args.probe_ckpt`
            # overrides the path; "" / "none" = explicit opt-out (random-init probe: synthetic benchmarks, tests);
            # a path that is given (or the reference's default when it exists) must load, a missing one warns loudly
            ckpt = getattr(args, "probe_ckpt", None)
            explicit = ckpt is not None
            if ckpt is None:
                ckpt = os.path.join(os.path.abspath("."), "models", "psdProbe_base_savel7.pt")
            if str(ckpt).lower() not in ("", "none"):
                if os.path.exists(ckpt):
                    self.oneWordpsdProbe.load_state_dict(
                        torch.load(ckpt, map_location="cpu", weights_only=False).state_dict())
                elif explicit:
                    raise FileNotFoundError("TVNetSAModel2: probe checkpoint %s not found" % ckpt)
                else:
                    import warnings
                    warnings.warn("TVNetSAModel2: %s not found -- the psdProbe matrix stays randomly initialised "
                                  "(the reference loads it unconditionally, models/bert_model.py:474-475); set "
                                  "args.probe_ckpt to the file, or to '' to silence this" % ckpt)
            self.combineLoss = CombineLoss(args.beta)
        self._engine: Optional[Engine] = None
        self._compute_dtype = _resolve_dtype(getattr(args, "compute_dtype", None))
        self.bert._engine_owner = self.engine
        self._last_heads = None

    # ------------------------------------------------------------------
    def set_compute_dtype(self, dtype):
        self._compute_dtype = _resolve_dtype(dtype)
        if self._engine is not None:
            self._engine.compute_dtype = self._compute_dtype

    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self, self.bert.hot_config(), "bert.", self._compute_dtype)
        return self._engine

    def _features(self, images, aux_imgs):
        """[n_img, B, 4, 3840] fp32 rows exactly as :538-539 build them (cat over the 4 pyramid levels,
        then a plain view(bsz, prefix_len, -1))."""
        with torch.no_grad():
            if getattr(self.image_model, "packed_features", False) and images.is_cuda:
                # features arrive packed (FeatureStub / the feature cache): one kernel does cat + view + aux permute +
                # cast to the GEMM dtype (fp32 or bf16 wire format in)
                if aux_imgs is not None and aux_imgs.dtype != images.dtype:
                    aux_imgs = aux_imgs.to(images.dtype)
                eng = self.engine()
                if not images[0].is_contiguous():
                    images = images.contiguous()
                if aux_imgs is not None and not aux_imgs[0].is_contiguous():
                    aux_imgs = aux_imgs.contiguous()
                out = ops.pack_features(images, aux_imgs, eng.compute_dtype)
                return out.view(out.shape[0], out.shape[1], self.args.prefix_len, -1)
            main, aux = self.image_model(images, aux_imgs)
            B = images.size(0)
            feats = [torch.cat(main, dim=1).reshape(B, self.args.prefix_len, -1)]
            if aux is not None:
                feats += [torch.cat(a, dim=1).reshape(B, self.args.prefix_len, -1) for a in aux]
            return torch.stack(feats).to(F32).contiguous()

    def get_visual_prompt(self, images, aux_imgs, imagelabel=None):
        """models/bert_model.py:534-588.  Returns (prefix, img_tag_loss, aux_img_tag_loss list); `prefix`
        is the packed [n_layers,2,B,P*H] tensor (index [l][0]/[l][1] and reshape(B,heads,P,d) give the
        reference's per-layer (key, value)) which `self.bert(past_key_values=prefix)` consumes as is."""
        eng = self.engine()
        eng.prepare()
        feats = self._features(images, aux_imgs)
        vao = bool(getattr(self.args, "vao", False))
        il = None
        if vao:
            il = imagelabel.to(device=feats.device, dtype=F32).contiguous()
        kv, img_losses = _FusionFn.apply(eng, feats, il, vao, self.training, 3, _grad_anchor(self.fc.weight))
        self._img_losses = img_losses if vao else None
        if vao:
            return kv, img_losses[0], [img_losses[j] for j in range(1, img_losses.numel())]
        return kv, 0, []

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, labels=None, imagelabel=None,
                images=None, aux_imgs=None):
        """models/bert_model.py:480-532."""
        if not input_ids.is_cuda:
            raise L.MtvafError("mtvaf_b200 runs on CUDA devices only (no CPU fallback); got %s" % input_ids.device)
        eng = self.engine()
        eng.new_step()
        eng.prepare()
        eng._nested = True
        try:
            return self._forward_impl(eng, input_ids, attention_mask, token_type_ids, labels, imagelabel, images,
                                      aux_imgs)
        finally:
            eng._nested = False

    def _forward_impl(self, eng, input_ids, attention_mask, token_type_ids, labels, imagelabel, images, aux_imgs):
        a = self.args
        B, Lq = input_ids.shape
        img_losses = None
        kv = None
        if a.use_prefix:
            kv, _, _ = self.get_visual_prompt(images, aux_imgs, imagelabel)
            img_losses = self._img_losses
            eng._kv_internal_ptr = kv.data_ptr()              # consumed once, by the encoder call right below
        out = self.bert(input_ids=input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids,
                        past_key_values=kv, output_attentions=True, output_hidden_states=True, return_dict=True)
        hs = out["hidden_states"]
        nl = self.bert.config.num_hidden_layers
        use_probe = bool(getattr(a, "use_probe", False))
        mask = attention_mask.to(torch.long).contiguous()
        lab = None if labels is None else labels.to(torch.long).contiguous()
        loss, prob = _HeadsFn.apply(eng, self, hs[min(7, nl)] if use_probe else None, hs[nl], mask, lab, img_losses,
                                    self.training)
        heads = self._last_heads
        logits = _DecodedTags(heads["best"], heads["lens"])
        self.last_emissions = heads["emissions"]
        alpha = float(getattr(a, "alpha", 0.0))
        img_term = alpha * (img_losses[:1].sum() if getattr(a, "noauxloss", False) else img_losses.sum()) \
            if img_losses is not None else torch.zeros((), device=input_ids.device)
        res = TokenClassifierOutput(loss=loss if labels is not None else None, logits=logits)
        if use_probe:
            return res, prob, img_term
        return res


class _SpanHeadsFn(torch.autograd.Function):
    """Heads + losses of the span variant (models/bert_model.py:288-312, 323-376) as one autograd node."""

    @staticmethod
    def forward(ctx, engine: Engine, model, hs_probe, hs_last, mask, batch, training):
        a = model.args
        B, Lq = mask.shape
        H = engine.cfg.H
        n = engine.cfg.n_layers
        probe_layer = min(7, n)
        hs = {n: hs_last.reshape(B * Lq, H), probe_layer: None}
        if hs_probe is not None:
            hs[probe_layer] = hs_probe.reshape(B * Lq, H)
        use_probe = bool(getattr(a, "use_probe", False)) and hs_probe is not None
        need_grad = any(ctx.needs_input_grad) and batch.get("polarity_labels") is not None
        out, saved = engine.span_heads_fwd(hs, B, Lq, mask, batch, use_probe, float(getattr(a, "beta", 0.5)),
                                           int(getattr(a, "num_epochs", 30)), training, need_grad,
                                           probe_layer=probe_layer)
        ctx.engine, ctx.saved = engine, saved
        ctx.shape = hs_last.shape
        ctx.has_probe_in = hs_probe is not None
        model._last_heads = out
        dev = mask.device
        loss = out["loss"] if out["loss"] is not None else torch.zeros(1, dtype=F32, device=dev)
        prob = out["prob_loss"] if out["prob_loss"] is not None else torch.zeros(1, dtype=F32, device=dev)
        tot = out.get("tot_loss")
        tot = tot if tot is not None else torch.zeros(1, dtype=F32, device=dev)
        ctx.mark_non_differentiable(prob, tot)
        return loss.view(()), prob.view(()), tot.view(())

    @staticmethod
    def backward(ctx, dloss, _dprob, _dtot):
        eng = ctx.engine
        if ctx.saved is None:
            raise L.MtvafError("backward through the span heads needs the training labels (no loss was computed)")
        eng.flat.attach_grads()
        grads = eng.span_heads_bwd(ctx.saved, dloss.reshape(1).to(F32).contiguous())
        n = eng.cfg.n_layers
        d_last = grads[n].view(ctx.shape)
        pl = ctx.saved["probe_layer"]
        d_probe = grads[pl].view(ctx.shape) if (ctx.has_probe_in and pl in grads and pl != n) else None
        ctx.saved = None
        return (None, None, d_probe, d_last, None, None, None)


class TVNetSAModel(nn.Module):
    """Span variant (models/bert_model.py:192-414; `--dataset_name twitter15|twitter17`, MTVAF_training.py:32-50):
    same encoder + visual prefix (no ANP heads), span extractor (`binary_affine`), span-pooled polarity classifier
    (`unary_affine`, `dense`, `classifier`), distant / mean cross-entropy losses.  GCN extras
    (`gcn_layer_number`, `num_layers`) and Cutoff augmentation are out of scope."""

    def __init__(self, label_list, tokenizer, args, type_num=None, use_weight=False, config=None,
                 image_model: Optional[nn.Module] = None):
        super().__init__()
        if getattr(args, "gcn_layer_number", 0) > 0 or getattr(args, "num_layers", 0) > 0:
            raise L.MtvafError("mtvaf_b200.TVNetSAModel: the GCN extras are outside the hot path (SURVEY.md 2, row 6)")
        self.args = args
        self.type_num = type_num
        self.tokenizer = tokenizer
        self.prefix_dim = args.prefix_dim
        self.prefix_len = args.prefix_len
        enc_cls = RobertaModel if "roberta" in args.bert_name else BertModel
        self.bert = enc_cls.from_config(config) if config is not None else enc_cls.from_pretrained(args.bert_name)
        H = self.bert.config.hidden_size
        nl = self.bert.config.num_hidden_layers
        # registration order of models/bert_model.py:205-240 (checkpoints are walked by index, train.py:495-521)
        self.dense = nn.Linear(H, H)
        self.activation = nn.Tanh()
        self.unary_affine = nn.Linear(H, 1)
        self.binary_affine = nn.Linear(H, 2)
        self.num_labels = len(label_list) + 1
        self.classifier = nn.Linear(H, 4)
        if args.use_prefix:
            if image_model is not None:
                self.image_model = image_model
            else:
                self.image_model = ImageModel(use_152=getattr(args, "use_152", False),
                                              use_101=getattr(args, "use_101", False),
                                              use_34=getattr(args, "use_34", False),
                                              use_18=getattr(args, "use_18", False), resnet_root=args.resnet_root)
            self.encoder_conv = nn.Sequential(nn.Linear(3840, 800), nn.Tanh(), nn.Linear(800, 4 * 2 * H))
            self.projectors = nn.ModuleList([nn.Linear(4 * H * 2, 4) for _ in range(nl)])
        self.fc = nn.Linear(H, self.num_labels)          # registered by the reference (:236), unused on this path
        self.dropout = nn.Dropout(0.1)
        if args.use_probe:
            self.oneWordpsdProbe = probe(args={"probe": {"maximum_rank": H // 2}, "model": {"hidden_dim": H}})
            self.combineLoss = CombineLoss(args.beta)
        self._engine: Optional[Engine] = None
        self._compute_dtype = _resolve_dtype(getattr(args, "compute_dtype", None))
        self.bert._engine_owner = self.engine
        self._last_heads = None

    def set_compute_dtype(self, dtype):
        self._compute_dtype = _resolve_dtype(dtype)
        if self._engine is not None:
            self._engine.compute_dtype = self._compute_dtype

    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self, self.bert.hot_config(), "bert.", self._compute_dtype)
        return self._engine

    _features = TVNetSAModel2._features

    def get_visual_prompt(self, images, aux_imgs):
        """models/bert_model.py:379-414 (the gate stack without the ANP heads); returns the packed prefix."""
        eng = self.engine()
        eng.prepare()
        feats = self._features(images, aux_imgs)
        kv, _ = _FusionFn.apply(eng, feats, None, False, self.training, 3, _grad_anchor(self.dense.weight))
        return kv

    def _encode(self, input_ids, attention_mask, token_type_ids, prefix_guids):
        return self.bert(input_ids=input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids,
                         past_key_values=prefix_guids, output_attentions=True, output_hidden_states=True,
                         return_dict=True)

    def extraction(self, prompt_attention_mask, input_ids, prefix_guids, token_type_ids, augument=False, labels=None,
                   adj_matrix=None, src_mask=None, aspect_mask=None):
        """models/bert_model.py:323-361: span proposals (start / end logits) for the trainer (modules/train.py:362-380).
        `prompt_attention_mask` is [B, P+L] as in the reference; the text part is its last L columns.

        Default: inference form -- the trainer discards this pass's autograd graph anyway (it detaches the logits,
        modules/train.py:384-386) and calls forward() for the loss, i.e. the reference runs the encoder twice per
        step.  With `args.reuse_extraction = True` (opt-in) this pass KEEPS its graph and its
        activations, and the forward() that follows on the same inputs continues from them instead of recomputing
        the visual prompt and the encoder: one encoder forward per step instead of two.  The only behavioural
        difference: both passes then share one dropout draw instead of two independent ones."""
        if augument:
            raise L.MtvafError("Cutoff augmentation is outside the hot path (SURVEY.md 2, row 12)")
        eng = self.engine()
        eng.prepare()
        B, Lq = input_ids.shape
        text_mask = prompt_attention_mask[:, -Lq:].to(torch.long).contiguous()
        reuse = bool(getattr(self.args, "reuse_extraction", False))
        self._stash = None
        with torch.set_grad_enabled(reuse and torch.is_grad_enabled()):
            out = self._encode(input_ids, text_mask, token_type_ids, prefix_guids)
            hs = out["hidden_states"]
        if reuse:
            self._stash = dict(key=self._stash_key(input_ids, token_type_ids, text_mask), hs=hs)
        with torch.no_grad():
            nl = self.bert.config.num_hidden_layers
            use_probe = bool(getattr(self.args, "use_probe", False))
            H = eng.cfg.H
            hsd = {nl: hs[nl].detach().reshape(B * Lq, H), min(7, nl): hs[min(7, nl)].detach().reshape(B * Lq, H)}
            o, _ = eng.span_heads_fwd(hsd, B, Lq, text_mask, {}, use_probe, float(getattr(self.args, "beta", 0.5)),
                                      int(getattr(self.args, "num_epochs", 30)), self.training, False,
                                      probe_layer=min(7, nl))
        seq = o["sequence_output"].view(B, Lq, H)
        if use_probe:
            return o["start_logits"], o["end_logits"], seq, o["prob_loss"].view(())
        return o["start_logits"], o["end_logits"], seq

    @staticmethod
    def _stash_key(input_ids, token_type_ids, text_mask):
        """Identity of an encoder input: same storage, same contents version, same values of the (small) mask."""
        tt = None if token_type_ids is None else (token_type_ids.data_ptr(), token_type_ids._version)
        return (input_ids.data_ptr(), input_ids._version, tuple(input_ids.shape), tt, tuple(text_mask.shape))

    def classification(self, span_starts, span_ends, sequence_input, attention_mask):
        """models/bert_model.py:363-376, inference form: (logits [B,M,4], ac_logits [B*M,4])."""
        eng = self.engine()
        eng.prepare()
        B, Lq, H = sequence_input.shape
        f = eng.flat
        with torch.no_grad():
            seq = sequence_input.reshape(B * Lq, H)
            seq32 = (ops.cast_f32(seq.contiguous()) if seq.dtype == BF16 else seq.to(F32)).contiguous()
            mask = attention_mask.to(torch.long).contiguous()
            ws = ops.span_offsets(mask)
            pooled = ops.span_pool_fwd(seq32, ws, span_starts.contiguous(), span_ends.contiguous(),
                                       f.w("unary_affine.weight").view(-1), f.w("unary_affine.bias"), B, Lq)
            pc = ops.cast_bf16(pooled) if eng.bf16 else pooled
            h = ops.linear_fwd(pc, eng.cw("dense.weight"), f.w("dense.bias"), mode=L.EPI_TANH)
            h32 = ops.cast_f32(h) if h.dtype == BF16 else h
            ac = ops.skinny_linear(h32, f.w("classifier.weight"), f.w("classifier.bias"))
        return ac.view(B, span_starts.shape[1], -1), ac

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, start_positions=None,
                end_positions=None, span_starts=None, span_ends=None, polarity_labels=None, label_masks=None,
                images=None, aux_imgs=None, valid_ids=None, adjacency_matrix=None, output_attention=False,
                augument=False, labels=None, adj_matrix=None, src_mask=None, aspect_mask=None, polaritys=None):
        """models/bert_model.py:246-321.  `label_masks` is accepted and -- exactly as in the reference, where the
        already-averaged cross-entropy is multiplied by the mask and divided by its sum (:302-303) -- has no
        effect on the loss."""
        if not input_ids.is_cuda:
            raise L.MtvafError("mtvaf_b200 runs on CUDA devices only (no CPU fallback); got %s" % input_ids.device)
        if augument:
            raise L.MtvafError("Cutoff augmentation is outside the hot path (SURVEY.md 2, row 12)")
        eng = self.engine()
        eng.new_step()
        eng.prepare()
        eng._nested = True
        try:
            a = self.args
            mask = attention_mask.to(torch.long).contiguous()
            stash, self._stash = getattr(self, "_stash", None), None
            if stash is not None and stash["key"] == self._stash_key(input_ids, token_type_ids, mask) and \
                    (stash["hs"][0].requires_grad or not torch.is_grad_enabled()):
                # args.reuse_extraction: continue from the activations (and autograd graph) of the extraction() pass
                hs = stash["hs"]
            else:
                kv = self.get_visual_prompt(images, aux_imgs) if a.use_prefix else None
                if kv is not None:
                    eng._kv_internal_ptr = kv.data_ptr()      # consumed once, by the encoder call right below
                hs = self._encode(input_ids, mask, token_type_ids, kv)["hidden_states"]
            nl = self.bert.config.num_hidden_layers
            use_probe = bool(getattr(a, "use_probe", False))
            lab = lambda t: None if t is None else t.to(torch.long).contiguous()
            batch = dict(start_positions=lab(start_positions), end_positions=lab(end_positions),
                         span_starts=lab(span_starts), span_ends=lab(span_ends), polarity_labels=lab(polarity_labels))
            loss, prob, tot = _SpanHeadsFn.apply(eng, self, hs[min(7, nl)] if use_probe else None, hs[nl], mask, batch,
                                                 self.training)
            heads = self._last_heads
            has_loss = heads["loss"] is not None
            res = TokenClassifierOutput(loss=loss if has_loss else None, logits=heads.get("logits"))
            if use_probe:
                return res, prob, tot
            return res
        finally:
            eng._nested = False


class _DecodedTags(list):
    """CRF decode result: behaves as the reference's List[List[int]] (models/bert_model.py:511) but the
    device->host copy happens only when the list is first read (no sync in the training step)."""

    def __init__(self, best: torch.Tensor, lens: torch.Tensor):
        super().__init__()
        self._best, self._lens, self._done = best, lens, False

    def _materialize(self):
        if not self._done:
            best, lens = self._best.cpu(), self._lens.cpu()
            list.extend(self, [best[b, :int(lens[b])].tolist() for b in range(best.shape[0])])
            self._done = True

    def __iter__(self):
        self._materialize()
        return list.__iter__(self)

    def __len__(self):
        self._materialize()
        return list.__len__(self)

    def __getitem__(self, i):
        self._materialize()
        return list.__getitem__(self, i)

    def __eq__(self, other):
        self._materialize()
        return list.__eq__(self, other)

    def __repr__(self):
        self._materialize()
        return list.__repr__(self)

    @property
    def device_tags(self):
        return self._best, self._lens
