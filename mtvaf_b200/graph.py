"""Whole-step CUDA graph: forward + backward + gradient all-reduce + AdamW of a drop-in model captured ONCE
and replayed per step.

The eager path issues ~450 kernel launches per training step from Python (ctypes -> C ABI); once the kernels
are fast the step is bounded by that host work.  A CUDA graph removes it: every launch of the step, the NCCL
all-reduces on the side stream and the optimizer are replayed by the driver from one `cudaGraphLaunch`.

What makes the step replayable (nothing that changes per step may be a kernel ARGUMENT):
  * dropout: every dropout site mixes a device-resident step counter into its seed
    (`mtvaf_set_step_source` / `mtvaf_advance_step`), advanced by the first node of the graph;
  * AdamW: step count, learning-rate schedule factor and bias corrections live in device memory
    (`FlatAdamW.enable_device_clock`, `mtvaf_adam_dyn_advance`);
  * inputs: the graph reads static device buffers; `__call__` copies the batch into them, `prefetch` stages the
    NEXT batch host->device on a copy stream while the current replay runs.
The reference has no counterpart (eager PyTorch, modules/train.py:859-885 `_step` + `optimizer.step()`); the
reference-facing nn.Module API keeps working eagerly next to this.

Construction does NOT train: the eager warm-up steps that prime the allocator pools / NCCL before the capture run on
`example_batch`, and weights, Adam moments, gradient buffer, the optimizer clock (host `t` and device `dyn`) and the
dropout step counter are snapshotted before and restored after it.  Step 1 of training is the first `__call__`.
"""
from __future__ import annotations

import weakref
from typing import Dict, Optional

import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer, example_batch: Dict[str, torch.Tensor], grad_sync=None,
                 warmup: int = 3):
        self.model, self.opt, self.sync = model, optimizer, grad_sync
        dev = next(p for p in model.parameters() if p.is_cuda).device
        self.device = dev
        self.static = {k: v.to(dev).clone() for k, v in example_batch.items()}
        self.staging: Optional[Dict[str, torch.Tensor]] = None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.staged_event: Optional[torch.cuda.Event] = None
        self.consumed_event: Optional[torch.cuda.Event] = None
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        ops.set_step_source(self.step_dev)
        # the registration is a process-global raw pointer: make sure it never outlives the tensor it points at
        self._finalizer = weakref.finalize(self, ops.clear_step_source_if, self.step_dev.data_ptr())
        optimizer.enable_device_clock()
        eng = optimizer.engine
        eng.prepare()
        snap = self._snapshot(eng, optimizer)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):              # eager warm-up off the default stream (torch.cuda.graphs rule)
            for _ in range(max(1, warmup)):
                self._step()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(self.graph, stream=side):      # same stream as the warm-up: no stale node streams
            self.loss = self._step()
        self.kernels_per_replay = ops.launch_count() - n0     # library kernels captured (NCCL / copies not counted)
        self.replays = 0
        # dropout seeds are baked into the captured launches from the host counter as of the capture (the device
        # counter `step_dev` is what varies per replay); tests replay the same masks eagerly from this value
        self.captured_host_step = eng.step_counter
        self._restore(eng, optimizer, snap)

    @staticmethod
    def _snapshot(eng, opt):
        f = eng.flat
        return dict(W=f.W.clone(), G=f.G.clone(), m=opt.m.clone(), v=opt.v.clone(), dyn=opt.dyn.clone(), t=opt.t,
                    host_step=eng.step_counter)

    def _restore(self, eng, opt, snap):
        f = eng.flat
        torch.cuda.synchronize(self.device)
        f.W.copy_(snap["W"])
        f.G.copy_(snap["G"])
        opt.m.copy_(snap["m"])
        opt.v.copy_(snap["v"])
        opt.dyn.copy_(snap["dyn"])
        opt.t = snap["t"]
        eng.step_counter = snap["host_step"]
        self.step_dev.zero_()
        if eng.bf16:
            f.refresh_bf16(force=True)
        torch.cuda.synchronize(self.device)

    def _step(self) -> torch.Tensor:
        ops.advance_step(self.step_dev)
        out = self.model(**self.static)
        loss = out[0].loss if isinstance(out, tuple) else out.loss
        loss.backward()
        # data parallel: the tail all-reduce is launched here and AdamW of the encoder layers runs beneath it
        self.opt.step(zero_grad=True, sync=self.sync)
        return loss.detach()

    # ------------------------------------------------------------------ inputs
    def prefetch(self, host_batch: Dict[str, torch.Tensor]):
        """Stage the NEXT batch (pinned host tensors) host->device on the copy stream; overlaps the running replay."""
        if self.staging is None:
            self.staging = {k: torch.empty_like(v) for k, v in self.static.items()}
        with torch.cuda.stream(self.copy_stream):
            if self.consumed_event is not None:    # the previous staged batch must have been consumed
                self.copy_stream.wait_event(self.consumed_event)
            for k, v in host_batch.items():
                self.staging[k].copy_(v, non_blocking=True)
            self.staged_event = torch.cuda.Event()
            self.staged_event.record(self.copy_stream)

    def __call__(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """One training step.  `batch` = device (or pinned host) tensors copied into the static inputs, or None to
        consume the batch staged by `prefetch`.  Returns the loss as a static device tensor (valid until the next
        call)."""
        cur = torch.cuda.current_stream(self.device)
        if batch is None:
            if self.staged_event is None:
                raise RuntimeError("GraphedTrainStep: no batch given and none staged by prefetch()")
            cur.wait_event(self.staged_event)
            for k, v in self.staging.items():
                self.static[k].copy_(v, non_blocking=True)
            self.consumed_event = torch.cuda.Event()
            self.consumed_event.record(cur)
        else:
            for k, v in batch.items():
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.loss

    def close(self):
        """Release the captured graph (and the NCCL kernels it references) and unregister the step counter."""
        self._finalizer()
        if self.graph is not None:
            torch.cuda.synchronize(self.device)
            self.graph.reset()
            self.graph = None
