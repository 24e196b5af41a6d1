"""Fused AdamW over the flat parameter buffer + per-rank gradient synchronisation.

Replaces `torch.optim.AdamW` with the reference's parameter groups (modules/train.py:894-926: names
containing 'bert' and 'encoder_conv' at args.lr, 'crf'/'fc*' at 5e-2, weight decay 1e-2, everything
else -- projectors, ANP heads, probe -- never updated), the linear warm-up/decay schedule
(`get_linear_schedule_with_warmup`, modules/train.py:118-120,919-921) and the broken
`modules/parallel.py` path: one process per GPU, NCCL all-reduce of the flat gradient buffer issued per
encoder layer from inside backward so it overlaps the remaining backward kernels.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops
from .engine import Engine, BF16, F32


def reference_groups(lr: float):
    """(predicate on parameter name, lr, weight_decay) exactly as modules/train.py:896-916."""
    return [(lambda n: "bert" in n, lr, 1e-2),
            (lambda n: "encoder_conv" in n or "gates" in n, lr, 1e-2),
            (lambda n: "crf" in n or n.startswith("fc"), 5e-2, 1e-2)]


class FlatAdamW:
    def __init__(self, engine: Engine, lr: float = 5e-5, betas=(0.9, 0.999), eps: float = 1e-8, groups=None,
                 warmup_steps: int = 0, total_steps: int = 0):
        self.engine = engine
        engine.prepare()
        f = engine.flat
        self.betas, self.eps = betas, eps
        self.warmup_steps, self.total_steps = warmup_steps, total_steps
        groups = groups if groups is not None else reference_groups(lr)
        # contiguous [start, end) ranges of the flat buffer sharing (lr, wd)
        self.ranges: List[Tuple[int, int, float, float]] = []
        cur = None
        for n in f.names:
            p = f.params[n]
            hit = None
            if p.requires_grad and n not in f.no_grad_names:
                for pred, glr, gwd in groups:
                    if pred(n):
                        hit = (glr, gwd)
                        break
            o, k = f.offsets[n]
            if hit is None:
                cur = None
                continue
            if cur is not None and tuple(cur[2:]) == hit and o - cur[1] < 64:
                cur[1] = o + k
            else:
                cur = [o, o + k, hit[0], hit[1]]
                self.ranges.append(cur)
        # a range must lie entirely inside or outside the bf16-shadowed prefix [0, cast_end) of the flat buffer: the
        # fused kernel refreshes the shadow of the ranges it updates, and a straddling range would silently keep a stale
        # shadow for its first part (latent with today's parameter order)
        split = []
        for a, b, lr_, wd_ in self.ranges:
            if a < f.cast_end < b:
                split += [[a, f.cast_end, lr_, wd_], [f.cast_end, b, lr_, wd_]]
            else:
                split.append([a, b, lr_, wd_])
        self.ranges = split
        self.m = torch.zeros(f.total, dtype=F32, device=f.device)
        self.v = torch.zeros(f.total, dtype=F32, device=f.device)
        self.t = 0
        self.dyn: Optional[torch.Tensor] = None    # device clock {uint64 t; float lr_scale, bc1, bc2_sqrt}

    def enable_device_clock(self):
        """Keep the step count, the learning-rate schedule factor and the Adam bias corrections in device memory
        (advanced by one tiny kernel per step) so that `step()` issues identical launches every time -- the form
        a CUDA graph can replay (mtvaf_b200.graph.GraphedTrainStep)."""
        if self.dyn is None:
            f = self.engine.flat
            self.dyn = torch.zeros(3, dtype=torch.int64, device=f.device)
            self.dyn[0] = self.t                   # steps taken so far (floats are rewritten by the advance kernel)
        return self.dyn

    def lr_scale(self) -> float:
        """get_linear_schedule_with_warmup evaluated for the step about to be taken."""
        if self.total_steps <= 0:
            return 1.0
        s = self.t
        if s < self.warmup_steps:
            return s / max(1.0, float(self.warmup_steps))
        return max(0.0, (self.total_steps - s) / max(1.0, float(self.total_steps - self.warmup_steps)))

    def step(self, grad_scale: float = 1.0, zero_grad: bool = False, sync: Optional["GradSync"] = None):
        """One AdamW update of every range.  zero_grad=True clears the whole flat gradient buffer in the same
        pass (the update kernels clear the ranges they own, one memset per gap of never-updated parameters),
        replacing `optimizer.zero_grad()` / `model.zero_grad()` and its one-fill-per-parameter launches.
        With `sync` (data parallel) the tail all-reduce is launched first and the encoder-layer ranges -- already
        reduced during backward -- are updated while it is in flight; the tail ranges follow once it lands."""
        f = self.engine.flat
        tail_lo = None
        if sync is not None and sync.world > 1:
            sync.launch_tail()
            sync.wait_layers()
            tail_lo = f.layer_ranges[-1][1] if f.layer_ranges else 0
        scale = self.lr_scale()
        self.t += 1
        if self.dyn is not None:
            # device clock: t, lr factor and bias corrections are advanced on the GPU; kernel arguments stay constant
            ops.adam_dyn_advance(self.dyn, self.betas[0], self.betas[1], self.warmup_steps, self.total_steps)
            scale = 1.0
        bf = self.engine.bf16 and f.Wb is not None
        pending = tail_lo is not None
        for a, b, lr, wd in self.ranges:
            if pending and (b > tail_lo or a < (f.layer_ranges[0][0] if f.layer_ranges else 0)):
                sync.wait_tail()            # ranges are sorted: everything from here on is tail
                pending = False
            shadow = None
            if bf and b <= f.cast_end:
                shadow = f.Wb[a:b]
            ops.adamw_step(f.W[a:b], f.G[a:b], self.m[a:b], self.v[a:b], lr * scale, self.betas[0], self.betas[1],
                           self.eps, wd, self.t, grad_scale, shadow, zero_grad, self.dyn)
        if pending:
            sync.wait_tail()
        if zero_grad:
            for a, b in self._gaps():
                f.G[a:b].zero_()
        if bf:
            f._wb_version = f.weights_version()     # shadow written by the fused kernel: still in sync

    def _gaps(self):
        if not hasattr(self, "_gap_list"):
            gaps, pos = [], 0
            for a, b, _, _ in sorted(self.ranges):
                if a > pos:
                    gaps.append((pos, a))
                pos = max(pos, b)
            if pos < self.engine.flat.total:
                gaps.append((pos, self.engine.flat.total))
            self._gap_list = gaps
        return self._gap_list

    def zero_grad(self):
        self.engine.flat.G.zero_()


class GradSync:
    """NCCL gradient all-reduce (average) of the flat gradient buffer, overlapped with backward.

    * per encoder layer, issued from inside backward on a side stream the moment layer i's gradients are final
      (`Engine.layer_grad_hook`) on the process group `group` -- bench.py caps that communicator at a few CTAs
      because these transfers have the rest of backward to hide behind;
    * the TAIL (heads, fusion MLP, embedding tables: final only when backward ends, so nothing hides it) goes out
      on `tail_group` (a second communicator WITHOUT the CTA cap when bench.py provides one), and
      `FlatAdamW.step(sync=...)` updates the already-reduced encoder layers while the tail is in flight;
    * `reserve_sms`: while layer all-reduces are resident the persistent kernels (static tile schedules) size their
      grids for `SMs - reserve_sms`, so none of their CTAs queues behind the SMs the collective holds;
    * with `optimizer=` only the ranges that optimizer updates are reduced: gradients of parameters the reference
      never steps (ANP heads, gate projectors, probe -- modules/train.py:894-926; 51 M of them) stay rank-local.

    Usage: sync = GradSync(engine, optimizer=opt); loss.backward(); opt.step(zero_grad=True, sync=sync)
       or: sync = GradSync(engine); loss.backward(); sync.finish(); optimizer.step()"""

    def __init__(self, engine: Engine, group=None, tail_group=None, optimizer: Optional["FlatAdamW"] = None,
                 reserve_sms: int = 0, tail_reserve_sms: int = 0):
        import torch.distributed as dist
        self.dist = dist
        self.engine = engine
        self.group = group
        self.tail_group = tail_group if tail_group is not None else group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.works = []            # layer all-reduces in flight
        self.tail_works = []
        self.done_layers = set()
        self.side = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.tail_side = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.owned = None if optimizer is None else sorted((a, b) for a, b, _, _ in optimizer.ranges)
        # SMs the persistent kernels leave alone while layer all-reduces are resident (mtvaf_set_sm_reserve): set
        # when the first layer hook fires, cleared in wait_layers()
        self.reserve_sms = reserve_sms
        self.tail_reserve_sms = tail_reserve_sms
        self._reserved = False
        self._emb_lo = None        # start of the embedding tables in the flat buffer (they are packed last)
        self._emb_done = False
        flat = engine.flat
        emb = [n for n in getattr(flat, "names", []) if ".embeddings." in n or n.startswith("embeddings.")]
        if emb and max(flat.offsets[n][0] + flat.offsets[n][1] for n in emb) + 64 >= flat.total:
            self._emb_lo = min(flat.offsets[n][0] for n in emb)
        self._enabled = True
        if self.world > 1:
            engine.layer_grad_hook = self._on_layer
            engine.tail_grad_hook = self._on_embeddings
            engine.pre_backward_hook = self._before_backward

    # ------------------------------------------------------------------ gradient accumulation
    def no_sync(self):
        """Context manager for the micro-batches of a gradient-accumulation window that do NOT end in an optimizer
        step (the reference trainer's `gradient_accumulation_steps`, modules/train.py:620): their backward passes only
        accumulate into the flat gradient buffer; the backward of the LAST micro-batch (outside this context) reduces
        the sums.  (Reducing every micro-batch is also correct -- an average of averages -- it only costs bandwidth.)"""
        sync = self

        class _Ctx:
            def __enter__(self_):
                sync._enabled = False

            def __exit__(self_, *exc):
                sync._enabled = True
                return False
        return _Ctx()

    def _before_backward(self):
        """A backward that starts while all-reduces of a previous backward are still in flight (two backward passes
        without an optimizer step between them) would accumulate into slices NCCL is still reading / writing on the
        side stream: make this stream wait for them first."""
        if self.works or self.tail_works:
            for w in self.works:
                w.wait()
            for w in self.tail_works:
                w.wait()
            self.works.clear()
            self.tail_works.clear()
            self.done_layers.clear()
            self._emb_done = False

    # ------------------------------------------------------------------ ranges
    def _clip(self, lo: int, hi: int):
        """[lo, hi) restricted to what must be reduced (everything, or the optimizer's ranges)."""
        if hi <= lo:
            return []
        if self.owned is None:
            return [(lo, hi)]
        out = []
        for a, b in self.owned:
            a2, b2 = max(a, lo), min(b, hi)
            if b2 > a2:
                if out and a2 - out[-1][1] < 1024:      # bridge alignment padding: fewer, larger collectives
                    out[-1] = (out[-1][0], b2)
                else:
                    out.append((a2, b2))
        return out

    def tail_ranges(self):
        f = self.engine.flat
        lo = f.layer_ranges[-1][1] if f.layer_ranges else 0
        first = f.layer_ranges[0][0] if f.layer_ranges else 0
        hi = self._emb_lo if (self._emb_done and self._emb_lo is not None) else f.total
        return self._clip(0, first) + self._clip(lo, hi)

    def _on_embeddings(self):
        """Called by Engine.encoder_bwd right after the embedding backward: the embedding tables are 80 % of the
        tail, and the fusion backward that still follows hides their all-reduce."""
        if self._emb_lo is None or self._emb_done or not self._enabled:
            return
        for a, b in self._clip(self._emb_lo, self.engine.flat.total):
            self._reduce(self.engine.flat.G[a:b], tail=True)
        self._emb_done = True
        if self.tail_reserve_sms and self.engine.flat.G.is_cuda:
            ops.set_sm_reserve(max(self.tail_reserve_sms, self.reserve_sms))   # wide collective resident from here on
            self._reserved = True

    def _reduce(self, t: torch.Tensor, tail: bool = False):
        group = self.tail_group if tail else self.group
        if not t.is_cuda:
            # host tensors (gloo; the CPU tests of this bookkeeping): synchronous SUM then 1/world
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=group)
            t.div_(self.world)
            return
        side = self.tail_side if tail else self.side
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(side):
            side.wait_event(ev)
            w = self.dist.all_reduce(t, op=self.dist.ReduceOp.AVG, group=group, async_op=True)
        (self.tail_works if tail else self.works).append(w)

    def _on_layer(self, i: int):
        if not self._enabled:
            return
        a, b = self.engine.flat.layer_ranges[i]
        self.done_layers.add(i)
        for a2, b2 in self._clip(a, b):
            self._reduce(self.engine.flat.G[a2:b2])
        if self.reserve_sms and not self._reserved and self.engine.flat.G.is_cuda:
            ops.set_sm_reserve(self.reserve_sms)
            self._reserved = True

    # ------------------------------------------------------------------ end of backward
    def launch_tail(self):
        """Issue what backward could not: layers whose hook never fired, then the tail ranges."""
        if self.world <= 1:
            return
        f = self.engine.flat
        for i, (a, b) in enumerate(f.layer_ranges):
            if i not in self.done_layers:
                for a2, b2 in self._clip(a, b):
                    self._reduce(f.G[a2:b2])
        for a, b in self.tail_ranges():
            self._reduce(f.G[a:b], tail=True)

    def wait_layers(self):
        if self._reserved:
            ops.set_sm_reserve(0)
            self._reserved = False
        for w in self.works:
            w.wait()
        self.works.clear()
        self.done_layers.clear()

    def wait_tail(self):
        for w in self.tail_works:
            w.wait()
        self.tail_works.clear()
        self._emb_done = False

    def finish(self):
        """Reduce everything backward has not reduced yet, then wait for all of it."""
        if self.world <= 1:
            return
        self.launch_tail()
        self.wait_layers()
        self.wait_tail()
