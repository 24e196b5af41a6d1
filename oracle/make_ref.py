"""TEST / BASELINE INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference where the GPU box can see it.

`/root/reference` exists only in the authoring container.  This recipe copies its Python sources and the two
shipped probe checkpoints, byte for byte, into `oracle/_ref/` (git-ignored, so nothing of the reference enters the
history; NOT gpurun-ignored, so the copy travels to the GPU box with the snapshot like our own built `.so`).  It is
used there as the CHECKER and as the BASELINE only:

  * `bench.py --impl reference`           times the unmodified reference modules on the host cores;
  * `bench.py` `gpu_eager_baseline`       times the same modules in PyTorch eager on the B200 (fp32 / bf16);
  * `tests/test_reference_trainer_gpu.py` drives the reference's own `SATrainer2` over the drop-in model;
  * `tests/test_oracle_vs_reference.py`   re-runs the oracle-vs-reference comparison on the GPU box.

Nothing under `mtvaf_b200/` imports it.  `__graft_entry__.build()` runs this when `/root/reference` is present.
    python -m oracle.make_ref
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MTVAF_REF_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP_EXT = (".py", ".pt")          # sources + probes/psdProbe_base_savel{4,7}.pt; no images, no README


def stage(verbose: bool = False) -> str:
    if not os.path.isdir(os.path.join(SRC, "models")):
        raise RuntimeError("reference tree not present at %s" % SRC)
    manifest = []
    for root, dirs, files in os.walk(SRC):
        dirs[:] = sorted(d for d in dirs if not d.startswith(".") and d != "__pycache__")
        for f in sorted(files):
            if not f.endswith(KEEP_EXT):
                continue
            s = os.path.join(root, f)
            rel = os.path.relpath(s, SRC)
            d = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            data = open(s, "rb").read()
            if not (os.path.exists(d) and open(d, "rb").read() == data):
                shutil.copyfile(s, d)
            manifest.append("%s  %s" % (hashlib.sha256(data).hexdigest(), rel))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    if verbose:
        print("staged %d reference files into %s" % (len(manifest), DST))
    return DST


if __name__ == "__main__":
    stage(verbose=True)
    sys.exit(0)
