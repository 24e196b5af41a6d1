"""Frozen visual front-end as an OFFLINE stage + the feature wire format.

The reference runs the frozen ResNet (`ImageModel`, models/bert_model.py:63-111) on the full image and the three object
crops of every sample in every training step (models/bert_model.py:536, modules/train.py:485-486,920-921 freeze it).
Its outputs are the INPUTS of the hot path: four pyramid maps average-pooled to 2 x 2, concatenated to
`[3840, 2, 2]` per image.  Since the front-end never trains, this module computes them once and stores them:

  wire format   one row per sample: `[1 + n_aux, 3840, 2, 2]` bf16 (image 0 = the full image, 1.. = the aux crops),
                61 440 values = 120 KiB per sample; a flat little-endian file with a fixed 4 KiB JSON header.
                bf16 is what the fusion GEMM consumes anyway (`mtvaf_pack_features` reads fp32 or bf16).
  FeatureCache  `build()` runs the torchvision ResNet (library conv kernels: out of scope for the hand-written path)
                over an iterable of image batches and appends rows; `open()` memory-maps the file; `batch(indices)`
                gathers rows into a PINNED staging buffer, `to_device()` issues the async H2D copy and returns
                `(images [B,3840,2,2], aux_imgs [B,n_aux,3840,2,2])` in bf16 -- the tensors `TVNetSAModel(2).forward`
                takes when its `image_model` is a `FeatureStub`.

Caveat: the reference's `image_process` applies RandomCrop / RandomHorizontalFlip at load time
(models/utils.py:593-600), so caching features freezes ONE augmentation draw per sample; pass `n_views > 1` to `build`
to store several draws and pick one at random per epoch.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

MAGIC = "MTVAF-FEATURES-1"
HEADER_BYTES = 4096
PYRAMID = 3840 * 2 * 2          # values per image


def pyramid_features(image_model: torch.nn.Module, images: torch.Tensor,
                     aux_imgs: Optional[torch.Tensor]) -> torch.Tensor:
    """[B, 1 + n_aux, 3840, 2, 2] fp32 from raw images through `ImageModel` (models/bert_model.py:88-111 semantics: the
    four pyramid levels of each image concatenated on the channel axis)."""
    with torch.no_grad():
        main, aux = image_model(images, aux_imgs)
        rows = [torch.cat(main, dim=1)]
        if aux is not None:
            rows += [torch.cat(a, dim=1) for a in aux]
        return torch.stack(rows, dim=1).float()


class FeatureCache:
    def __init__(self, path: str, n_samples: int, n_aux: int, mm: np.memmap):
        self.path, self.n_samples, self.n_aux, self._mm = path, n_samples, n_aux, mm
        self.row = (1 + n_aux) * PYRAMID
        self._pinned: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ writing
    @classmethod
    def build(cls, path: str, image_model: torch.nn.Module, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]],
              n_samples: int, n_aux: int = 3, device: Optional[torch.device] = None) -> "FeatureCache":
        """batches: iterable of (images [b,3,H,W], aux_imgs [b,n_aux,3,H,W]) in dataset order."""
        row = (1 + n_aux) * PYRAMID
        header = json.dumps({"magic": MAGIC, "n_samples": n_samples, "n_aux": n_aux, "row_values": row,
                             "dtype": "bfloat16", "layout": "[n_samples, 1+n_aux, 3840, 2, 2]"}).encode()
        assert len(header) < HEADER_BYTES
        with open(path, "wb") as fh:
            fh.write(header.ljust(HEADER_BYTES, b"\0"))
            fh.truncate(HEADER_BYTES + 2 * n_samples * row)
        mm = np.memmap(path, dtype=np.uint16, mode="r+", offset=HEADER_BYTES, shape=(n_samples, row))
        image_model = image_model.eval()
        if device is not None:
            image_model = image_model.to(device)
        at = 0
        for images, aux in batches:
            if device is not None:
                images, aux = images.to(device), (None if aux is None else aux.to(device))
            f = pyramid_features(image_model, images, aux)                       # [b, 1+n_aux, 3840, 2, 2]
            f16 = f.to(torch.bfloat16).reshape(f.shape[0], row).cpu()
            mm[at:at + f16.shape[0]] = f16.view(torch.int16).numpy().view(np.uint16)
            at += f16.shape[0]
        assert at == n_samples, "wrote %d of %d samples" % (at, n_samples)
        mm.flush()
        return cls.open(path)

    @classmethod
    def from_tensor(cls, path: str, feats: torch.Tensor) -> "FeatureCache":
        """feats [N, 1 + n_aux, 3840, 2, 2] (any float dtype) -> file."""
        n, n_img = feats.shape[0], feats.shape[1]
        row = n_img * PYRAMID
        header = json.dumps({"magic": MAGIC, "n_samples": n, "n_aux": n_img - 1, "row_values": row,
                             "dtype": "bfloat16", "layout": "[n_samples, 1+n_aux, 3840, 2, 2]"}).encode()
        with open(path, "wb") as fh:
            fh.write(header.ljust(HEADER_BYTES, b"\0"))
            fh.write(feats.to(torch.bfloat16).reshape(n, row).contiguous().cpu().view(torch.int16).numpy().tobytes())
        return cls.open(path)

    # ------------------------------------------------------------------ reading
    @classmethod
    def open(cls, path: str) -> "FeatureCache":
        with open(path, "rb") as fh:
            raw = fh.read(HEADER_BYTES)
        try:
            h = json.loads(raw.rstrip(b"\0").decode())
        except Exception as e:
            raise ValueError("%s is not a feature cache (bad header)" % path) from e
        if h.get("magic") != MAGIC or h.get("dtype") != "bfloat16":
            raise ValueError("%s is not a %s file" % (path, MAGIC))
        n, n_aux = int(h["n_samples"]), int(h["n_aux"])
        row = (1 + n_aux) * PYRAMID
        if int(h["row_values"]) != row or os.path.getsize(path) != HEADER_BYTES + 2 * n * row:
            raise ValueError("%s: size does not match its header" % path)
        mm = np.memmap(path, dtype=np.uint16, mode="r", offset=HEADER_BYTES, shape=(n, row))
        return cls(path, n, n_aux, mm)

    def __len__(self):
        return self.n_samples

    def batch(self, indices: Sequence[int], pinned: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Rows `indices` gathered into a pinned host tensor [B, 1 + n_aux, 3840, 2, 2] bf16 (reused between calls
        unless `pinned` is given: double-buffer it when copies overlap the next gather)."""
        B = len(indices)
        buf = pinned
        if buf is None:
            if self._pinned is None or self._pinned.shape[0] < B:
                self._pinned = torch.empty((B, self.row), dtype=torch.bfloat16)
                if torch.cuda.is_available():
                    self._pinned = self._pinned.pin_memory()
            buf = self._pinned
        dst = buf[:B].view(torch.int16).numpy().view(np.uint16)
        np.take(self._mm, np.asarray(indices, dtype=np.int64), axis=0, out=dst)
        return buf[:B].view(B, 1 + self.n_aux, 3840, 2, 2)

    @staticmethod
    def to_device(host_rows: torch.Tensor, device, stream: Optional["torch.cuda.Stream"] = None):
        """Async H2D of one gathered batch; returns (images, aux_imgs) bf16 views of ONE device buffer."""
        if stream is not None:
            with torch.cuda.stream(stream):
                d = host_rows.to(device, non_blocking=True)
        else:
            d = host_rows.to(device, non_blocking=True)
        return d[:, 0], (d[:, 1:] if d.shape[1] > 1 else None)
