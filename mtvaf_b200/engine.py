"""Host-side orchestration of the hot path: parameter packing + functional forward/backward sequences.

Every arithmetic step below is a call into the C ABI (mtvaf_b200.ops); this module only decides the
order of launches and owns the saved-activation bookkeeping, i.e. it is the hand-written replacement
of PyTorch autograd for the fixed MTVAF graph:

  fusion_fwd/bwd   TVNetSAModel2.get_visual_prompt        models/bert_model.py:534-588
  encoder_fwd/bwd  RobertaModel/BertModel.forward         models/modeling_roberta.py:850-978, :480-568
  heads_fwd/bwd    TVNetSAModel2.forward tail             models/bert_model.py:503-527, probes/*

Data layout in HBM (per model):
  * all float parameters live in ONE flat fp32 buffer `W` (nn.Parameters are views into it); per
    encoder layer the q/k/v weights (and biases) are adjacent so the fused QKV projection reads one
    [3H,H] matrix; gradients live in a matching flat fp32 buffer `G` (param.grad are views) that the
    wgrad kernels accumulate into directly (split-K fp32 atomics) and NCCL all-reduces per layer slice;
  * bf16 mode keeps a bf16 shadow `Wb` of the GEMM weights, refreshed by one cast launch per step;
  * activations are [T=B*L, features] row-major in the compute dtype.
"""
from __future__ import annotations

import itertools
from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import lib as L
from . import ops

F32 = torch.float32
BF16 = torch.bfloat16
_ALIGN = 64   # elements (256 B): keeps every view 16-byte aligned for TMA / vector access


class HotPathConfig:
    def __init__(self, kind: str, hidden: int, heads: int, inter: int, layers: int, eps: float, pad_id: int,
                 hidden_dropout: float = 0.1, attn_dropout: float = 0.1):
        self.kind, self.H, self.nh, self.I, self.n_layers = kind, hidden, heads, inter, layers
        self.eps, self.pad_id = eps, pad_id
        self.d = hidden // heads
        self.hidden_dropout, self.attn_dropout = hidden_dropout, attn_dropout
        if self.d != 64:
            raise L.MtvafError("mtvaf_b200 attention kernels support head_dim 64 only (got %d)" % self.d)
        if hidden % 64 != 0:
            raise L.MtvafError("hidden size must be a multiple of 64")


_LAYER_ORDER = ["attention.self.query.weight", "attention.self.key.weight", "attention.self.value.weight",
                "attention.self.query.bias", "attention.self.key.bias", "attention.self.value.bias",
                "attention.output.dense.weight", "attention.output.dense.bias",
                "attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias",
                "intermediate.dense.weight", "intermediate.dense.bias",
                "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"]


class FlatParams:
    """Packs the float parameters of a module into flat fp32 weight / gradient buffers (+ bf16 shadow)."""

    def __init__(self, module: torch.nn.Module, enc_prefix: str, n_layers: int):
        self.module = module
        self.enc_prefix = enc_prefix
        # the frozen ResNet front-end (out of scope) keeps its own storage
        named = OrderedDict((n, p) for n, p in module.named_parameters()
                            if p.dtype == F32 and not n.startswith("image_model."))
        order: List[str] = []
        self.layer_ranges: List[Tuple[int, int]] = []
        layer_names: List[List[str]] = []
        for i in range(n_layers):
            names = ["%sencoder.layer.%d.%s" % (enc_prefix, i, s) for s in _LAYER_ORDER]
            if all(n in named for n in names):
                layer_names.append(names)
                order += names
        emb = [n for n in named if n.startswith(enc_prefix + "embeddings.")]
        rest = [n for n in named if n not in set(order) and n not in set(emb)]
        # the gate projectors (models/bert_model.py:455: one Linear(8H, 4) per layer) are consumed as ONE [4n, 8H] matrix
        # (+ [4n] bias) by the single gate GEMM: pack their weights, then their biases, contiguously -- no torch.cat
        # per forward, and their gradients are views of the flat buffer the wgrad kernel accumulates into directly
        pj_w = [n for n in rest if n.startswith("projectors.") and n.endswith(".weight")]
        pj_b = [n for n in rest if n.startswith("projectors.") and n.endswith(".bias")]
        pj_w.sort(key=lambda n: int(n.split(".")[1]))
        pj_b.sort(key=lambda n: int(n.split(".")[1]))
        self.projector_names = (pj_w, pj_b)
        rest = [n for n in rest if n not in set(pj_w) and n not in set(pj_b)] + pj_w + pj_b
        order += rest
        self.n_cast_names = len(order)           # everything before the embedding tables gets a bf16 shadow
        order += emb
        self.names = order
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = 0
        qkv_w = {s for s in _LAYER_ORDER[:3]}
        qkv_b = {s for s in _LAYER_ORDER[3:6]}
        cast_end = 0
        for idx, n in enumerate(order):
            p = named[n]
            tail = n.split("encoder.layer.")[-1].split(".", 1)[-1] if "encoder.layer." in n else ""
            packed = tail in qkv_w or tail in qkv_b      # q,k,v blocks must be exactly adjacent
            follows = packed and tail not in (_LAYER_ORDER[0], _LAYER_ORDER[3])
            follows = follows or (n in pj_w[1:]) or (n in pj_b[1:])
            if not follows:
                off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
            self.offsets[n] = (off, p.numel())
            off += p.numel()
            if idx == self.n_cast_names - 1:
                cast_end = off
        self.total = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        self.cast_end = (cast_end + 7) // 8 * 8
        for names in layer_names:
            a = self.offsets[names[0]][0]
            b = self.offsets[names[-1]][0] + self.offsets[names[-1]][1]
            self.layer_ranges.append((a, b))
        # the pooler is dead on the MTVAF path: like the reference it gets NO gradient
        self.no_grad_names = {n for n in order if ".pooler." in n or n.startswith("pooler.")}
        self.device = None
        self.W = self.G = self.Wb = None
        self._wb_version = -1
        self.params = named

    # ------------------------------------------------------------------
    def _is_attached(self) -> bool:
        if self.W is None:
            return False
        n0 = self.names[0]
        p = self.params[n0]
        return p.data_ptr() == self.W.data_ptr() + 4 * self.offsets[n0][0] and p.device == self.W.device

    def ensure(self, device: torch.device):
        """(Re)attach after construction / .to(device) / load_state_dict with assign."""
        if self._is_attached():
            return
        first = next(iter(self.params.values()))
        if not first.is_cuda:
            raise L.MtvafError("mtvaf_b200: model parameters are on %s -- move the model to a CUDA device; "
                               "the hot path has no CPU implementation" % first.device)
        device = first.device
        W = torch.zeros(self.total, dtype=F32, device=device)
        for n in self.names:
            p = self.params[n]
            o, k = self.offsets[n]
            W[o:o + k].copy_(p.data.reshape(-1))
        self.W = W
        self.G = torch.zeros(self.total, dtype=F32, device=device)
        self.Wb = None
        self._wb_version = -1
        for n in self.names:
            p = self.params[n]
            o, k = self.offsets[n]
            p.data = W[o:o + k].view(p.shape)
            p.grad = None
        self.device = device

    def w(self, name: str) -> torch.Tensor:
        o, k = self.offsets[name]
        return self.W[o:o + k].view(self.params[name].shape)

    def g(self, name: str) -> torch.Tensor:
        o, k = self.offsets[name]
        return self.G[o:o + k].view(self.params[name].shape)

    def wb(self, name: str) -> torch.Tensor:
        o, k = self.offsets[name]
        return self.Wb[o:o + k].view(self.params[name].shape)

    def span(self, buf: torch.Tensor, first: str, last: str, shape) -> torch.Tensor:
        a = self.offsets[first][0]
        b = self.offsets[last][0] + self.offsets[last][1]
        return buf[a:b].view(shape)

    def weights_version(self) -> int:
        return sum(p._version for p in self.params.values())

    def refresh_bf16(self, force: bool = False):
        v = self.weights_version()
        if self.Wb is None:
            self.Wb = torch.empty(self.total, dtype=BF16, device=self.device)
            force = True
        if force or v != self._wb_version:
            ops.cast_bf16(self.W[:self.cast_end], self.Wb[:self.cast_end])
            self._wb_version = v

    def attach_grads(self) -> bool:
        """Point param.grad at the flat gradient buffer. Returns True if this starts a fresh
        accumulation (grads were None -> buffer zeroed), False if accumulating into existing grads."""
        if not hasattr(self, "_live"):
            self._live = [n for n in self.names if self.params[n].requires_grad and n not in self.no_grad_names]
        live = self._live
        # fast path (called from every autograd node of the step): first and last live params already attached
        p0, p1 = self.params[live[0]], self.params[live[-1]]
        if (p0.grad is not None and p1.grad is not None and p0.grad.data_ptr() == self.g(live[0]).data_ptr()
                and p1.grad.data_ptr() == self.g(live[-1]).data_ptr()):
            return False
        fresh = any(self.params[n].grad is None for n in live)
        if fresh:
            self.G.zero_()
        for n in live:
            p = self.params[n]
            if True:
                g = self.g(n)
                if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                    p.grad = g
        return fresh


# =================================================================================================
class Engine:
    """Launch sequences for one model instance (encoder + optional fusion / heads)."""

    def __init__(self, module: torch.nn.Module, cfg: HotPathConfig, enc_prefix: str = "bert.",
                 compute_dtype: torch.dtype = BF16):
        self.cfg = cfg
        self.enc_prefix = enc_prefix
        self.flat = FlatParams(module, enc_prefix, cfg.n_layers)
        self.compute_dtype = compute_dtype
        self.step_counter = 0
        # dropout streams follow torch.manual_seed and differ per data-parallel rank (masks are counter-based hashes
        # of (base_seed, step, site, element): nothing else distinguishes two ranks or two runs)
        rank = 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank = dist.get_rank()
        except Exception:
            rank = 0
        self.base_seed = ((torch.initial_seed() ^ 0x5EED) + 7919 * rank) & 0x7FFFFFFF
        self._kv_internal_ptr = None       # prefix tensor handed from the fusion stack to the encoder within one forward
        self._dkv32: Dict[int, torch.Tensor] = {}
        self.pre_backward_hook: Optional[Callable[[], None]] = None    # DP: called before a backward touches the flat G
        self.layer_grad_hook: Optional[Callable[[int], None]] = None   # DP: called when layer i's grads are final
        self.tail_grad_hook: Optional[Callable[[], None]] = None       # DP: called when the embedding grads are final

    # ------------------------------------------------------------------ helpers
    @property
    def bf16(self) -> bool:
        return self.compute_dtype == BF16

    def prepare(self):
        if getattr(self, "_nested", False):      # already prepared by the enclosing model forward
            return
        dev = next(iter(self.flat.params.values())).device
        self.flat.ensure(dev)
        if self.bf16:
            self.flat.refresh_bf16()

    def _forked(self, n: int):
        """Context manager: `n` cached side streams that wait for the current stream on entry (fork) and that the current
        stream waits for on exit (join).  Tensors touched inside must outlive the join (they do: saved activations /
        function-scope buffers)."""
        eng = self
        if not hasattr(self, "_side_streams"):
            self._side_streams = []
        dev = self.flat.device

        class _Fork:
            def __enter__(self_):
                while len(eng._side_streams) < n:
                    eng._side_streams.append(torch.cuda.Stream(device=dev))
                self_.cur = torch.cuda.current_stream(dev)
                self_.streams = eng._side_streams[:n]
                for st in self_.streams:
                    st.wait_stream(self_.cur)
                return self_.streams

            def __exit__(self_, *exc):
                for st in self_.streams:
                    self_.cur.wait_stream(st)
                return False
        return _Fork()

    def new_step(self):
        """Advance the dropout step counter -- once per outermost forward entry point (a wrapping model's forward
        marks its nested encoder / fusion calls with `_nested`)."""
        if not getattr(self, "_nested", False):
            self.step_counter += 1

    def cw(self, name: str) -> torch.Tensor:
        """GEMM weight in the compute dtype."""
        return self.flat.wb(name) if self.bf16 else self.flat.w(name)

    def cspan(self, first: str, last: str, shape) -> torch.Tensor:
        return self.flat.span(self.flat.Wb if self.bf16 else self.flat.W, first, last, shape)

    def seed(self, site: int) -> int:
        return ((self.base_seed * 1000003 + self.step_counter) * 4099 + site) & 0xFFFFFFFFFFFFFFFF

    def lname(self, i: int, s: str) -> str:
        return "%sencoder.layer.%d.%s" % (self.enc_prefix, i, s)

    # ------------------------------------------------------------------ encoder
    def encoder_fwd(self, ids: torch.Tensor, tts: torch.Tensor, mask: torch.Tensor, kv: Optional[torch.Tensor],
                    training: bool, save: bool, want_probs: bool = False, embeds: Optional[torch.Tensor] = None):
        """ids/tts/mask [B,L] int64 on device; kv [n_layers,2,B,P*H] (compute dtype) or None.
        Returns (hidden_states list of n+1 [T,H] tensors, saved dict, attentions or None)."""
        c, f, e = self.cfg, self.flat, self.enc_prefix
        B, Lq = ids.shape
        H, nh, d = c.H, c.nh, c.d
        cd = self.compute_dtype
        p_h = c.hidden_dropout if training else 0.0
        p_a = c.attn_dropout if training else 0.0
        saved = {"B": B, "L": Lq, "ids": ids, "tts": tts, "mask": mask, "kv": kv, "p_h": p_h, "p_a": p_a,
                 "layers": [], "step": self.step_counter}
        if embeds is None:
            x, pids, mean, rstd = ops.embed_ln_fwd(ids, tts, f.w(e + "embeddings.word_embeddings.weight"),
                                                   f.w(e + "embeddings.position_embeddings.weight"),
                                                   f.w(e + "embeddings.token_type_embeddings.weight"),
                                                   f.w(e + "embeddings.LayerNorm.weight"),
                                                   f.w(e + "embeddings.LayerNorm.bias"), c.eps,
                                                   0 if c.kind == "roberta" else 1, c.pad_id, cd, p_h, self.seed(1))
            saved.update(pids=pids, emb_mean=mean, emb_rstd=rstd)
        else:
            x = embeds
            saved.update(pids=None)
        hs = [x]
        attns = [] if want_probs else None
        P = 0 if kv is None else kv.shape[3] // H
        for i in range(c.n_layers):
            ln = lambda s: self.lname(i, s)
            wqkv = self.cspan(ln(_LAYER_ORDER[0]), ln(_LAYER_ORDER[2]), (3 * H, H))
            bqkv = f.span(f.W, ln(_LAYER_ORDER[3]), ln(_LAYER_ORDER[5]), (3 * H,))
            qkv = ops.linear_fwd(x, wqkv, bqkv)
            kp = vp = None
            if kv is not None:
                kp = kv[i, 0].view(B, nh, P, d)
                vp = kv[i, 1].view(B, nh, P, d)
            ctx, lse, probs = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, nh, d, p_a, self.seed(16 * i + 2),
                                                want_probs)
            if want_probs:
                attns.append(probs)
            z1 = ops.linear_fwd(ctx, self.cw(ln("attention.output.dense.weight")),
                                f.w(ln("attention.output.dense.bias")), mode=L.EPI_RESID, aux=x, p_drop=p_h,
                                seed=self.seed(16 * i + 3))
            a, m1, r1 = ops.layernorm_fwd(z1, f.w(ln("attention.output.LayerNorm.weight")),
                                          f.w(ln("attention.output.LayerNorm.bias")), c.eps)
            # training: the forward epilogue also computes gelu'(pre) -- it has issue slots to spare, the backward
            # epilogue that would otherwise recompute it from the pre-activation does not (profiles/r2_ncu_hot_b512.md)
            # -- and `pre` holds that derivative instead of the pre-activation
            pre = torch.empty((B * Lq, c.I), dtype=cd, device=x.device) if save else None
            g = ops.linear_fwd(a, self.cw(ln("intermediate.dense.weight")), f.w(ln("intermediate.dense.bias")),
                               mode=L.EPI_GELU_GRAD if save else L.EPI_GELU, out2=pre)
            z2 = ops.linear_fwd(g, self.cw(ln("output.dense.weight")), f.w(ln("output.dense.bias")),
                                mode=L.EPI_RESID, aux=a, p_drop=p_h, seed=self.seed(16 * i + 4))
            y, m2, r2 = ops.layernorm_fwd(z2, f.w(ln("output.LayerNorm.weight")), f.w(ln("output.LayerNorm.bias")),
                                          c.eps)
            if save:
                saved["layers"].append(dict(x=x, qkv=qkv, ctx=ctx, lse=lse, z1=z1, m1=m1, r1=r1, a=a, pre=pre, g=g,
                                            z2=z2, m2=m2, r2=r2))
            x = y
            hs.append(x)
        return hs, saved, attns

    def encoder_bwd(self, saved, grad_hs: List[Optional[torch.Tensor]], want_dkv: bool,
                    want_dembeds: bool = False):
        """grad_hs[i]: gradient w.r.t. hidden_states[i] ([T,H], any float dtype) or None.
        Accumulates parameter gradients into the flat buffer; returns dkv fp32 [n_layers,2,B,P*H] or None
        (and the gradient w.r.t. the embedding output when the caller supplied `embeds`)."""
        c, f, e = self.cfg, self.flat, self.enc_prefix
        B, Lq, H, nh, d = saved["B"], saved["L"], c.H, c.nh, c.d
        cd = self.compute_dtype
        if self.pre_backward_hook:
            self.pre_backward_hook()
        kv, mask = saved["kv"], saved["mask"]
        p_h, p_a = saved["p_h"], saved["p_a"]
        step = saved["step"]
        sd = lambda site: ((self.base_seed * 1000003 + step) * 4099 + site) & 0xFFFFFFFFFFFFFFFF
        P = 0 if kv is None else kv.shape[3] // H
        # every layer that runs its attention backward overwrites its whole [2, B, P*H] slice (plain stores): only the
        # layers no gradient reaches are cleared (600 MB at B=512 otherwise memset per step)
        dkv = torch.empty((c.n_layers, 2, B, P * H), dtype=F32, device=mask.device) if (kv is not None and want_dkv) else None
        T = B * Lq

        def as_cd(t):
            t = t.reshape(T, H)
            if t.dtype != cd:
                t = ops.cast_bf16(t.contiguous()) if cd == BF16 else ops.cast_f32(t.contiguous())
            return t.contiguous()

        dy = None
        top = grad_hs[c.n_layers]
        if top is not None:
            dy = as_cd(top)
        for i in range(c.n_layers - 1, -1, -1):
            if dy is None:
                # no gradient reaches this layer's output (yet): only an injected grad can start the chain
                if grad_hs[i] is not None:
                    dy = as_cd(grad_hs[i])
                if dkv is not None:
                    dkv[i].zero_()
                if self.layer_grad_hook:
                    self.layer_grad_hook(i)
                continue
            s = saved["layers"][i]
            ln = lambda nm: self.lname(i, nm)
            # ---- output block: y = LN(dropout(g W2^T + b2) + a)
            dz2, dd2 = ops.layernorm_bwd(dy, s["z2"], f.w(ln("output.LayerNorm.weight")), s["m2"], s["r2"],
                                         f.g(ln("output.LayerNorm.weight")), f.g(ln("output.LayerNorm.bias")),
                                         d_bias=f.g(ln("output.dense.bias")), p_drop=p_h, seed=sd(16 * i + 4))
            ops.linear_wgrad(dd2, s["g"], f.g(ln("output.dense.weight")))
            # (d(intermediate.dense.bias) = column sums of dpre: summed from the staging boxes of this GEMM's epilogue)
            dpre = ops.linear_dgrad(dd2, self.cw(ln("output.dense.weight")), mode=L.EPI_MUL_AUX, aux=s["pre"],
                                    colsum=f.g(ln("intermediate.dense.bias")))
            # ---- intermediate: g = gelu(a W1^T + b1)
            ops.linear_wgrad(dpre, s["a"], f.g(ln("intermediate.dense.weight")))
            da = ops.linear_dgrad(dpre, self.cw(ln("intermediate.dense.weight")), mode=L.EPI_RESID, aux=dz2)
            # ---- attention output block: a = LN(dropout(ctx Wo^T + bo) + x)
            dz1, dd1 = ops.layernorm_bwd(da, s["z1"], f.w(ln("attention.output.LayerNorm.weight")), s["m1"], s["r1"],
                                         f.g(ln("attention.output.LayerNorm.weight")),
                                         f.g(ln("attention.output.LayerNorm.bias")),
                                         d_bias=f.g(ln("attention.output.dense.bias")), p_drop=p_h,
                                         seed=sd(16 * i + 3))
            ops.linear_wgrad(dd1, s["ctx"], f.g(ln("attention.output.dense.weight")))
            dctx = ops.linear_dgrad(dd1, self.cw(ln("attention.output.dense.weight")))
            # ---- attention core
            kp = vp = dkp = dvp = None
            if kv is not None:
                kp, vp = kv[i, 0].view(B, nh, P, d), kv[i, 1].view(B, nh, P, d)
                if dkv is not None:
                    dkp, dvp = dkv[i, 0], dkv[i, 1]
            # (the bias gradient of the fused QKV projection = column sums of dqkv comes out of the same call)
            dqkv = ops.attention_bwd(dctx, s["qkv"], kp, vp, mask, s["ctx"], s["lse"], B, Lq, nh, d, dkp, dvp, p_a,
                                     sd(16 * i + 2),
                                     d_bias=f.span(f.G, ln(_LAYER_ORDER[3]), ln(_LAYER_ORDER[5]), (3 * H,)))
            # ---- fused QKV projection
            wqkv = self.cspan(ln(_LAYER_ORDER[0]), ln(_LAYER_ORDER[2]), (3 * H, H))
            ops.linear_wgrad(dqkv, s["x"], f.span(f.G, ln(_LAYER_ORDER[0]), ln(_LAYER_ORDER[2]), (3 * H, H)))
            dx = ops.linear_dgrad(dqkv, wqkv, mode=L.EPI_RESID, aux=dz1)
            if grad_hs[i] is not None:
                ops.add_inplace(dx, grad_hs[i].reshape(T, H).contiguous())
            dy = dx
            saved["layers"][i] = None      # free activations as we go
            if self.layer_grad_hook:
                self.layer_grad_hook(i)
        d_embeds = None
        if dy is not None:
            if saved["pids"] is not None:
                ops.embed_ln_bwd(dy, saved["ids"], saved["tts"], saved["pids"],
                                 f.w(e + "embeddings.word_embeddings.weight"),
                                 f.w(e + "embeddings.position_embeddings.weight"),
                                 f.w(e + "embeddings.token_type_embeddings.weight"),
                                 f.w(e + "embeddings.LayerNorm.weight"), saved["emb_mean"], saved["emb_rstd"],
                                 0 if c.kind == "roberta" else 1, c.pad_id,
                                 f.g(e + "embeddings.word_embeddings.weight"),
                                 f.g(e + "embeddings.position_embeddings.weight"),
                                 f.g(e + "embeddings.token_type_embeddings.weight"),
                                 f.g(e + "embeddings.LayerNorm.weight"), f.g(e + "embeddings.LayerNorm.bias"),
                                 p_h, sd(1))
                if self.tail_grad_hook:
                    self.tail_grad_hook()      # DP: the embedding tables' gradients are final
            elif want_dembeds:
                d_embeds = dy
        return dkv, d_embeds

    # ------------------------------------------------------------------ fusion (visual prompt)
    def fusion_fwd(self, feats: torch.Tensor, imagelabel: Optional[torch.Tensor], vao: bool, training: bool,
                   save: bool, n_aux_heads: int):
        """feats [n_img, B, 4, 3840] pyramid rows, fp32 or already in the compute dtype (image 0 = full image, 1.. = aux
        crops).
        Returns (kv [n_layers,2,B,P*H] compute dtype, img_losses [n_img] fp32 or None, saved)."""
        c, f = self.cfg, self.flat
        cd = self.compute_dtype
        n_img, B = feats.shape[0], feats.shape[1]
        H = c.H
        W8 = 8 * H
        rows4 = n_img * B * 4
        x = feats.reshape(rows4, feats.shape[-1])
        if x.dtype == cd:
            x = x.contiguous()                      # packed by mtvaf_pack_features in the GEMM dtype already
        elif cd == BF16:
            x = ops.cast_bf16(x.contiguous())
        else:
            x = ops.cast_f32(x.contiguous())
        h1 = ops.linear_fwd(x, self.cw("encoder_conv.0.weight"), f.w("encoder_conv.0.bias"), mode=L.EPI_TANH)
        guids = ops.linear_fwd(h1, self.cw("encoder_conv.2.weight"), f.w("encoder_conv.2.bias"))     # [rows4, 8H]
        rows = n_img * B
        saved = dict(x=x, h1=h1, guids=guids, n_img=n_img, B=B, vao=vao)
        img_losses = None
        if vao:
            if n_img - 1 > n_aux_heads:
                raise L.MtvafError("vao supports at most %d aux images" % n_aux_heads)
            gm = ops.mean4_fwd(guids, rows, W8, 0)                                               # [rows, 8H]
            p_i = 0.2 if training else 0.0                                                       # img_dropout
            gmd = ops.dropout_apply(gm, p_i, self.seed(900)) if p_i > 0 else gm
            n_anp = f.params["img_classifier.weight"].shape[0]
            ld = (n_anp + 7) // 8 * 8
            logits = torch.zeros((rows, ld), dtype=F32, device=x.device)
            names = ["img_classifier"] + ["aux_img_classifier.%d" % k for k in range(n_img - 1)]
            # the ANP heads are independent M = B GEMMs that fill a quarter of the GPU each (18 CTA pairs at B = 512):
            # issued on side streams they run side by side (fork / join; captured as parallel branches of the graph)
            with self._forked(len(names)) as streams:
                for j, nm in enumerate(names):
                    with torch.cuda.stream(streams[j]):
                        ops.gemm(gmd[j * B:(j + 1) * B], self.cw(nm + ".weight"), M=B, N=n_anp, K=W8,
                                 bias=f.w(nm + ".bias"), out=logits[j * B:(j + 1) * B], ldo=ld)
            img_losses, dlogits = ops.softmax_kl(logits, n_anp, imagelabel, B, save)
            saved.update(gmd=gmd, dlogits=dlogits, names=names, n_anp=n_anp, p_i=p_i, seed_i=self.seed(900))
        gs = ops.mean4_fwd(guids, rows, W8, 1)                                                   # [rows, 8H]
        # all 12 projectors in one skinny GEMM: [rows, 8H] x [n_layers*4, 8H]^T
        pw, pb = self._projector_pack(f.W)
        if cd == BF16:
            # tensor cores, no split-K (deterministic): [rows, 8H] x [48, 8H]^T with fp32 logits out
            pwb, _ = self._projector_pack(f.Wb)
            gate_logits = ops.gemm(gs, pwb, M=rows, N=4 * c.n_layers, K=W8, bias=pb, out_dtype=F32)
        else:
            gate_logits = ops.skinny_linear(gs, pw, pb)          # [rows, 48], deterministic warp-per-row kernel
        kv, gates = ops.gate_fwd(guids, gate_logits, c.n_layers, n_img, B, H)
        saved.update(gs=gs, gate_logits=gate_logits, gates=gates)
        return kv, img_losses, (saved if save else None)

    def _projector_pack(self, buf: torch.Tensor):
        """projectors[l].weight [4, 8H] / bias [4] as ONE [4*n_layers, 8H] matrix (+ [4*n_layers] bias): views of the
        flat buffer `buf` (W, Wb or G) -- FlatParams packs them contiguously in layer order."""
        f, c = self.flat, self.cfg
        pj_w, pj_b = f.projector_names
        if len(pj_w) != c.n_layers:
            raise L.MtvafError("expected %d gate projectors, found %d" % (c.n_layers, len(pj_w)))
        return (f.span(buf, pj_w[0], pj_w[-1], (4 * c.n_layers, 8 * c.H)),
                f.span(buf, pj_b[0], pj_b[-1], (4 * c.n_layers,)))

    def fusion_bwd(self, saved, dkv: torch.Tensor, d_img_losses: Optional[torch.Tensor]):
        """dkv fp32 [n_layers,2,B,P*H]; d_img_losses: device fp32 [n_img] weights of the ANP losses."""
        c, f = self.cfg, self.flat
        n_img, B, H = saved["n_img"], saved["B"], c.H
        W8 = 8 * H
        rows, rows4 = n_img * B, n_img * B * 4
        cd = self.compute_dtype
        guids = saved["guids"]
        d_guids = torch.empty((rows4, W8), dtype=F32, device=guids.device)      # written by gate_bwd
        d_gate_logits = ops.gate_bwd(dkv, guids, saved["gate_logits"], saved["gates"], c.n_layers, n_img, B, H,
                                     d_guids)
        # projector GEMM backward: gradients accumulate straight into the packed [4n, 8H] / [4n] views of the flat
        # gradient buffer (no temporaries, no per-layer scatter)
        dpw, dpb = self._projector_pack(f.G)
        gs = saved["gs"]
        if cd == BF16:
            dgl = ops.cast_bf16(d_gate_logits)                                                   # [rows, 48]: tiny
            ops.linear_wgrad(dgl, gs, dpw)                                                       # tcgen05, K = rows
            pwb, _ = self._projector_pack(f.Wb)
            d_gs = ops.linear_dgrad(dgl, pwb, out_dtype=F32)                                     # [rows, 8H] fp32
        else:
            pw, _ = self._projector_pack(f.W)
            ops.linear_wgrad(d_gate_logits, gs, dpw)
            d_gs = ops.linear_dgrad(d_gate_logits, pw)                                           # [rows, 8H] fp32
        ops.colsum(d_gate_logits, dpb)
        d_gmd, p_i, seed_i = None, 0.0, 0
        if saved["vao"] and saved["dlogits"] is not None and d_img_losses is not None:
            dlog = saved["dlogits"]
            n_anp = saved["n_anp"]
            # scale each head's rows by the weight of its loss (device scalars, no sync)
            for j in range(n_img):
                ops.scale_by_device_scalar(dlog[j * B:(j + 1) * B], d_img_losses[j:j + 1])
            dl = ops.cast_bf16(dlog) if cd == BF16 else dlog
            d_gmd = torch.empty((rows, W8), dtype=cd, device=guids.device)
            with self._forked(len(saved["names"])) as streams:
                for j, nm in enumerate(saved["names"]):
                    with torch.cuda.stream(streams[j]):
                        dj = dl[j * B:(j + 1) * B]
                        ops.gemm(dj, self.cw(nm + ".weight"), b_mn=True, M=B, N=W8, K=n_anp,
                                 out=d_gmd[j * B:(j + 1) * B])
                        ops.linear_wgrad(dj, saved["gmd"][j * B:(j + 1) * B], f.g(nm + ".weight"), n_valid=n_anp)
                        ops.colsum(dj, f.g(nm + ".bias"), n_valid=n_anp)
            p_i, seed_i = saved["p_i"], saved["seed_i"]
        # gate path + both 4-way-mean backward terms (+ img_dropout mask) in one pass, in the GEMM dtype
        dg = ops.prompt_grad_combine(d_guids, d_gs, d_gmd, p_i, seed_i, rows, W8, cd)
        ops.linear_wgrad(dg, saved["h1"], f.g("encoder_conv.2.weight"))
        ops.colsum(dg, f.g("encoder_conv.2.bias"))
        dh1 = ops.linear_dgrad(dg, self.cw("encoder_conv.2.weight"), mode=L.EPI_MUL_DTANH, aux=saved["h1"])
        ops.linear_wgrad(dh1, saved["x"], f.g("encoder_conv.0.weight"))
        ops.colsum(dh1, f.g("encoder_conv.0.bias"))
        # inputs (frozen ResNet features) need no gradient

    # ------------------------------------------------------------------ heads
    def heads_fwd(self, hs: List[torch.Tensor], B: int, Lq: int, mask: torch.Tensor, labels: Optional[torch.Tensor],
                  use_probe: bool, beta: float, alpha: float, img_losses: Optional[torch.Tensor], training: bool,
                  save: bool, probe_layer: int = 7, epoch: int = 30):
        c, f = self.cfg, self.flat
        T, H = B * Lq, c.H
        seq = hs[c.n_layers]
        p_d = 0.1 if training else 0.0                                             # self.dropout, bert_model.py:466,506
        n_tags = f.params["fc.weight"].shape[0]
        if seq.dtype == BF16:
            # throughput mode: the tag head stays in bf16 on the tensor cores (fp32 accumulate, fp32 emissions).  The
            # fp32 detour (cast + fp32 dropout + warp-per-row kernels re-streaming fc.weight per row) cost 0.39 ms per
            # step for 0.55 GFLOP (profiles/r1_launches_v7): N = 11 wastes most of a 128-wide MMA tile, and is
            # still 5x faster because the pass is a single 50 MB read
            seq_d = ops.dropout_apply(seq, p_d, self.seed(910)) if p_d > 0 else seq
            em = ops.linear_fwd(seq_d, self.cw("fc.weight"), f.w("fc.bias"), out_dtype=F32).view(B, Lq, n_tags)
        else:
            seq_d = ops.dropout_apply(seq, p_d, self.seed(910)) if p_d > 0 else seq
            if n_tags <= 48 and H % 4 == 0:
                em = ops.skinny_linear(seq_d, f.w("fc.weight"), f.w("fc.bias")).view(B, Lq, n_tags)
            else:
                em = ops.linear_fwd(seq_d, f.w("fc.weight"), f.w("fc.bias")).view(B, Lq, n_tags)
        crf = (f.w("crf.start_transitions"), f.w("crf.end_transitions"), f.w("crf.transitions"))
        best, lens = ops.crf_decode(em, mask, *crf)
        out = dict(emissions=em, best=best, lens=lens, loss=None, prob_loss=None)
        saved = dict(B=B, L=Lq, seq_d=seq_d, p_d=p_d, seed_d=self.seed(910), use_probe=use_probe,
                     probe_layer=probe_layer)
        nll = None
        if labels is not None:
            nll, d_em, d_s, d_e, d_t = ops.crf_nll(em, labels, mask, *crf, save, 1.0 / B)
            saved.update(d_em=d_em, d_s=d_s, d_e=d_e, d_t=d_t)
        prob_loss = None
        if use_probe:
            x7 = hs[probe_layer]
            proj = f.params["oneWordpsdProbe.oneWordpsdProbe.proj"]
            r = proj.shape[1]
            projc = self.cw("oneWordpsdProbe.oneWordpsdProbe.proj")
            if x7.dtype == BF16 and r % 8 == 0:
                # T through the TMA-store epilogue, norms from it in one HBM-bound pass (the fused squared-norm epilogue
                # took 410 us for this 19 GFLOP GEMM: per-thread row stores + one fp32 atomic per 32 columns)
                Tm = ops.gemm(x7, projc, b_mn=True, M=T, N=r, K=H)
                norms = ops.row_sqnorm(Tm)
            else:
                norms = torch.zeros(T, dtype=F32, device=seq.device)
                Tm = torch.empty((T, r), dtype=x7.dtype, device=seq.device)
                ops.gemm(x7, projc, b_mn=True, M=T, N=r, K=H, mode=L.EPI_SQNORM, rowvec=norms, out=Tm)
            plabels = ops.probe_labels(norms.view(B, Lq))
            prob_loss, dnorms = ops.mse(norms, plabels.view(-1), save)
            out.update(prob_loss=prob_loss, norms=norms.view(B, Lq), pseudo_labels=plabels)
            saved.update(Tm=Tm, dnorms=dnorms, x7=x7)
        if nll is not None:
            loss, flag = ops.combine_loss(nll, B, prob_loss, beta, epoch, img_losses, alpha)
            out.update(loss=loss, crf_nll=nll)
            saved.update(flag=flag, probe_coef=beta * 2.0 ** (-epoch))
        return out, (saved if save else None)

    def heads_bwd(self, saved, dloss: torch.Tensor, dprob: Optional[torch.Tensor] = None):
        """dloss: device fp32 [1] = d(objective)/d(loss); dprob: device fp32 [1] = d(objective)/d(prob_loss) when
        someone differentiates the returned probe loss directly (None on the training path).  Returns grads w.r.t.
        hidden_states (dict index -> [T,H] fp32 tensor) and accumulates head parameter grads."""
        c, f = self.cfg, self.flat
        B, Lq, H = saved["B"], saved["L"], c.H
        T = B * Lq
        if self.pre_backward_hook:
            self.pre_backward_hook()
        grads = {}
        d_em = saved["d_em"]
        n_tags = d_em.shape[-1]
        for t in (d_em, saved["d_s"], saved["d_e"], saved["d_t"]):
            ops.scale_by_device_scalar(t, dloss)
        ops.add_inplace(f.g("crf.start_transitions"), saved["d_s"])
        ops.add_inplace(f.g("crf.end_transitions"), saved["d_e"])
        ops.add_inplace(f.g("crf.transitions"), saved["d_t"])
        de = d_em.view(T, n_tags)
        ops.colsum(de, f.g("fc.bias"))
        if saved["seq_d"].dtype == BF16:
            # bf16 mode: both gradients of the head on the tensor cores.  d(emissions) is padded to 16 columns (rows of
            # a bf16 operand must be 16-byte multiples); K = n_tags: the TMA boxes zero-fill past the 11 real tags
            de16 = torch.zeros((T, 16 * ((n_tags + 15) // 16)), dtype=BF16, device=de.device)
            de16[:, :n_tags].copy_(de)                                              # tiny (T x 16) cast, plumbing
            ops.gemm(de16, saved["seq_d"], a_mn=True, b_mn=True, M=n_tags, N=H, K=T, mode=L.EPI_ATOMIC_F32,
                     out=f.g("fc.weight"), splits=ops.wgrad_splits(n_tags, H, T, True))
            dseq = ops.gemm(de16, self.cw("fc.weight"), b_mn=True, M=T, N=H, K=n_tags)
            if saved["p_d"] > 0:
                dseq = ops.dropout_apply(dseq, saved["p_d"], saved["seed_d"])
        elif n_tags <= 16 and H % 8 == 0 and n_tags * (H // 8) <= 1536:
            # skinny head: dedicated kernels; the data gradient comes out dropout-masked in the encoder's dtype
            ops.skinny_linear_wgrad(de, saved["seq_d"], f.g("fc.weight"))
            dseq = ops.skinny_linear_dgrad(de, f.w("fc.weight"), self.compute_dtype, saved["p_d"], saved["seed_d"])
        else:
            ops.linear_wgrad(de, saved["seq_d"], f.g("fc.weight"))
            dseq = ops.linear_dgrad(de, f.w("fc.weight"))                           # [T,H] fp32
            if saved["p_d"] > 0:
                dseq = ops.dropout_apply(dseq, saved["p_d"], saved["seed_d"])
        grads[c.n_layers] = dseq
        if saved["use_probe"]:
            # loss += [prob_loss > 0.1] * prob_loss * beta * 2^-epoch  (probes/loss.py:14-16)
            dn = saved["dnorms"]
            flagf = saved["flag"].to(F32)          # 0/1 on device (tiny cast, plumbing)
            if dprob is None:
                ops.scale_by_device_scalar(dn, dloss)
                ops.scale_by_device_scalar(dn, flagf)
                alpha = 2.0 * saved["probe_coef"]
            else:
                # d/d(norms) of  dloss * [prob>0.1] * coef * prob  +  dprob * prob   (scalar glue on the device)
                ops.scale_by_device_scalar(dn, (dloss * flagf * saved["probe_coef"] + dprob).contiguous())
                alpha = 2.0
            Tm, x7 = saved["Tm"], saved["x7"]
            dT = ops.rowscale(Tm, dn, alpha)
            ops.linear_wgrad(x7, dT, f.g("oneWordpsdProbe.oneWordpsdProbe.proj"))
            projc = self.cw("oneWordpsdProbe.oneWordpsdProbe.proj")
            dx7 = ops.gemm(dT, projc, M=T, N=H, K=Tm.shape[1], out_dtype=self.compute_dtype)
            grads[saved["probe_layer"]] = dx7
        return grads


# =================================================================================================
# span variant heads: TVNetSAModel.extraction / classification / losses (models/bert_model.py:288-376)
# =================================================================================================
def _span_heads_fwd(self, hs: List[torch.Tensor], B: int, Lq: int, mask: torch.Tensor, batch: Dict[str, torch.Tensor],
                    use_probe: bool, beta: float, epoch: int, training: bool, save: bool, probe_layer: int = 7):
    """batch: start_positions / end_positions [B,L], span_starts / span_ends [B,M], polarity_labels [B,M] (int64)
    or None entries for inference.  Returns (out dict, saved)."""
    c, f = self.cfg, self.flat
    T, H = B * Lq, c.H
    cd = self.compute_dtype
    seq = hs[c.n_layers]
    seq32 = ops.cast_f32(seq) if seq.dtype == BF16 else seq
    p_d = 0.1 if training else 0.0                                             # self.dropout, bert_model.py:237,349
    seq_d = ops.dropout_apply(seq32, p_d, self.seed(910)) if p_d > 0 else seq32
    ae = ops.skinny_linear(seq_d, f.w("binary_affine.weight"), f.w("binary_affine.bias"))      # [T,2]  :351
    out = dict(ae=ae, start_logits=ae.view(B, Lq, 2)[..., 0], end_logits=ae.view(B, Lq, 2)[..., 1], loss=None,
               prob_loss=None, sequence_output=seq_d)
    saved = dict(B=B, L=Lq, seq_d=seq_d, p_d=p_d, seed_d=self.seed(910), use_probe=use_probe, probe_layer=probe_layer)
    starts, ends = batch.get("span_starts"), batch.get("span_ends")
    if starts is not None:
        M = starts.shape[1]
        ws = ops.span_offsets(mask)
        w_u = f.w("unary_affine.weight").view(-1)
        pooled = ops.span_pool_fwd(seq_d, ws, starts, ends, w_u, f.w("unary_affine.bias"), B, Lq)   # :364-369
        pooled_c = ops.cast_bf16(pooled) if cd == BF16 else pooled
        h_t = ops.linear_fwd(pooled_c, self.cw("dense.weight"), f.w("dense.bias"), mode=L.EPI_TANH)  # :371-372
        h_d = ops.dropout_apply(h_t, p_d, self.seed(911)) if p_d > 0 else h_t
        h32 = ops.cast_f32(h_d) if h_d.dtype == BF16 else h_d
        ac = ops.skinny_linear(h32, f.w("classifier.weight"), f.w("classifier.bias"))             # [B*M,4] :374
        out.update(ac_logits=ac, logits=ac.view(B, M, -1))
        saved.update(ws=ws, starts=starts, ends=ends, pooled_c=pooled_c, h_t=h_t, h32=h32, seed_h=self.seed(911), M=M)
    prob_loss = None
    if use_probe:
        x7 = hs[probe_layer]
        proj = f.params["oneWordpsdProbe.oneWordpsdProbe.proj"]
        r = proj.shape[1]
        projc = self.cw("oneWordpsdProbe.oneWordpsdProbe.proj")
        if x7.dtype == BF16 and r % 8 == 0:
            Tm = ops.gemm(x7, projc, b_mn=True, M=T, N=r, K=H)
            norms = ops.row_sqnorm(Tm)
        else:
            norms = torch.zeros(T, dtype=F32, device=seq.device)
            Tm = torch.empty((T, r), dtype=x7.dtype, device=seq.device)
            ops.gemm(x7, projc, b_mn=True, M=T, N=r, K=H, mode=L.EPI_SQNORM, rowvec=norms, out=Tm)
        plabels = ops.probe_labels(norms.view(B, Lq))
        prob_loss, dnorms = ops.mse(norms, plabels.view(-1), save)
        out.update(prob_loss=prob_loss, norms=norms.view(B, Lq), pseudo_labels=plabels)
        saved.update(Tm=Tm, dnorms=dnorms, x7=x7)
    sp, ep, pol = batch.get("start_positions"), batch.get("end_positions"), batch.get("polarity_labels")
    if sp is not None and ep is not None and pol is not None and starts is not None:
        tot = torch.zeros(1, dtype=F32, device=seq.device)
        d_ae = torch.empty_like(ae) if save else None
        ops.distant_ce(ae, 0, sp, 0.5, tot, d_ae)                               # (start_loss + end_loss) / 2  :298-300
        ops.distant_ce(ae, 1, ep, 0.5, tot, d_ae)
        d_ac = ops.ce_mean(out["ac_logits"], pol.reshape(-1).contiguous(), 1.0, tot, save)   # :302-303
        loss, flag = ops.combine_loss(tot, 1, prob_loss, beta, epoch, None, 0.0)             # :312
        out.update(loss=loss, tot_loss=tot)
        saved.update(d_ae=d_ae, d_ac=d_ac, flag=flag, probe_coef=beta * 2.0 ** (-epoch))
    return out, (saved if save else None)


def _span_heads_bwd(self, saved, dloss: torch.Tensor):
    c, f = self.cfg, self.flat
    B, Lq, H = saved["B"], saved["L"], c.H
    T = B * Lq
    cd = self.compute_dtype
    if self.pre_backward_hook:
        self.pre_backward_hook()
    grads = {}
    d_ae, d_ac = saved["d_ae"], saved["d_ac"]
    ops.scale_by_device_scalar(d_ae, dloss)
    ops.scale_by_device_scalar(d_ac, dloss)
    # classifier <- dropout <- tanh <- dense <- pooled
    ops.skinny_linear_wgrad(d_ac, saved["h32"], f.g("classifier.weight"))
    ops.colsum(d_ac, f.g("classifier.bias"))
    d_pre = ops.skinny_linear_dgrad(d_ac, f.w("classifier.weight"), cd, saved["p_d"], saved["seed_h"],
                                    tanh_out=saved["h_t"])                         # [B*M,H]
    ops.linear_wgrad(d_pre, saved["pooled_c"], f.g("dense.weight"))
    ops.colsum(d_pre, f.g("dense.bias"))
    d_pooled = ops.gemm(d_pre, self.cw("dense.weight"), b_mn=True, M=d_pre.shape[0], N=H, K=H, out_dtype=F32)
    # binary_affine, then the span gather scatters on top of it
    d_seq = ops.skinny_linear_dgrad(d_ae, f.w("binary_affine.weight"), F32)
    ops.skinny_linear_wgrad(d_ae, saved["seq_d"], f.g("binary_affine.weight"))
    ops.colsum(d_ae, f.g("binary_affine.bias"))
    ops.span_pool_bwd(d_pooled, saved["seq_d"], saved["ws"], saved["starts"], saved["ends"],
                      f.w("unary_affine.weight").view(-1), f.w("unary_affine.bias"), B, Lq, d_seq,
                      f.g("unary_affine.weight").view(-1), f.g("unary_affine.bias"))
    if saved["p_d"] > 0:
        d_seq = ops.dropout_apply(d_seq, saved["p_d"], saved["seed_d"])
    grads[c.n_layers] = ops.cast_bf16(d_seq) if cd == BF16 else d_seq
    if saved["use_probe"]:
        dn = saved["dnorms"]
        ops.scale_by_device_scalar(dn, dloss)
        ops.scale_by_device_scalar(dn, saved["flag"].to(F32))
        Tm, x7 = saved["Tm"], saved["x7"]
        dT = ops.rowscale(Tm, dn, 2.0 * saved["probe_coef"])
        ops.linear_wgrad(x7, dT, f.g("oneWordpsdProbe.oneWordpsdProbe.proj"))
        projc = self.cw("oneWordpsdProbe.oneWordpsdProbe.proj")
        grads[saved["probe_layer"]] = ops.gemm(dT, projc, M=T, N=H, K=Tm.shape[1], out_dtype=cd)
    return grads


Engine.span_heads_fwd = _span_heads_fwd
Engine.span_heads_bwd = _span_heads_bwd
