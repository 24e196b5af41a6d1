"""Host-sync removal around the hot path: the per-step bookkeeping of the reference's trainers without
per-step device synchronisation.

The reference's `SATrainer2.train` (modules/train.py:605-661) does, EVERY step: `loss.detach().cpu().item()` (three times
with the probe), `labels.to('cpu').numpy()`, `attention_mask.to('cpu').numpy()`, a B x L Python loop that indexes the
decoded tag lists (which forces the CRF decode's device->host copy) and appends label strings.  Each of those is a full
stream synchronisation in the middle of training.  The two helpers below keep the same RESULTS --

  * `TagLog`     collects (attention_mask, labels, decoded tags) per step as device tensors, copies them to pinned host
                 memory asynchronously, and builds `y_true` / `y_pred` (the seqeval inputs of modules/train.py:664)
                 ONCE per epoch with the reference's rules (skip column 0, stop at the first masked column, skip "X"
                 and "[SEP]" labels);
  * `LossMeter`  accumulates the step losses on the device and reads them back once per `refresh_step` window (the
                 reference prints window averages, modules/train.py:649-660)

-- with one synchronisation per refresh window / per epoch instead of several per step.  They are host-side Python by
design (the reference's trainer is; north_star keeps it): nothing here is on the kernel path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


class LossMeter:
    """Window averages of device scalars without a per-step `.item()`."""

    def __init__(self, names=("loss", "prob_loss", "img_loss"), refresh_step: int = 2):
        self.names = tuple(names)
        self.refresh_step = refresh_step
        self._acc: Optional[torch.Tensor] = None
        self._n = 0

    def add(self, **values) -> Optional[Dict[str, float]]:
        """values: name -> 0-d tensor (or python number).  Returns the window averages when a window closes (one D2H
        read), else None (no synchronisation)."""
        vals = []
        dev = None
        for n in self.names:
            v = values.get(n, 0.0)
            if torch.is_tensor(v):
                v = v.detach().reshape(()).float()
                dev = v.device
            vals.append(v)
        row = torch.stack([v if torch.is_tensor(v) else torch.tensor(float(v), device=dev) for v in vals])
        self._acc = row if self._acc is None else self._acc + row
        self._n += 1
        if self._n % self.refresh_step:
            return None
        out = (self._acc / self.refresh_step).cpu().tolist()          # the window's only synchronisation
        self._acc = None
        return dict(zip(self.names, out))


class TagLog:
    """Per-epoch `y_true` / `y_pred` of modules/train.py:627-647, built from device tensors copied out asynchronously."""

    def __init__(self, label_map: Dict[str, int]):
        # id -> label string, 0 = "PAD" (modules/train.py:629-630)
        self.id2label = {idx: label for label, idx in label_map.items()}
        self.id2label[0] = "PAD"
        self._pending: List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Optional[torch.cuda.Event]]] = []

    def append(self, attention_mask: torch.Tensor, labels: torch.Tensor, decoded) -> None:
        """decoded: the model's `logits` (mtvaf_b200 `_DecodedTags`: device tensors best [B, L] / lens [B]) or a
        plain List[List[int]] (the reference model).  No synchronisation."""
        if hasattr(decoded, "device_tags"):
            best, _ = decoded.device_tags
        else:
            Lq = labels.shape[1]
            best = torch.tensor([list(r) + [0] * (Lq - len(r)) for r in decoded], dtype=torch.long)
        host = []
        ev = None
        for t in (attention_mask, labels, best):
            t = t.detach()
            if t.is_cuda:
                h = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                h.copy_(t, non_blocking=True)
            else:
                h = t.clone()
            host.append(h)
        if attention_mask.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
        self._pending.append((host[0], host[1], host[2], ev))

    def finalize(self) -> Tuple[List[List[str]], List[List[str]]]:
        """Builds (y_true, y_pred) for everything appended since the last call; the epoch's only synchronisation."""
        y_true: List[List[str]] = []
        y_pred: List[List[str]] = []
        for mask, labels, best, ev in self._pending:
            if ev is not None:
                ev.synchronize()
            m, lab, pred = mask.numpy(), labels.numpy(), best.numpy()
            for row in range(m.shape[0]):
                t_row, p_row = [], []
                for col in range(1, m.shape[1]):                 # column 0 ([CLS]) is skipped (:637-638)
                    if not m[row, col]:
                        break                                    # stop at the first masked column (:645-646)
                    name = self.id2label[int(lab[row, col])]
                    if name != "X" and name != "[SEP]":
                        t_row.append(name)
                        p_row.append(self.id2label[int(pred[row, col])])
                y_true.append(t_row)
                y_pred.append(p_row)
        self._pending.clear()
        return y_true, y_pred
