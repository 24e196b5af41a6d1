"""In-tree build of the CUDA extension: explicit nvcc for sm_100a -> mtvaf_b200/_C/libmtvaf_b200.so.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  No JIT cache, no
torch.utils.cpp_extension: the library has a plain C ABI (include/mtvaf_b200.h) and links only cudart.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB_PATH = os.path.join(OUT_DIR, "libmtvaf_b200.so")
# experiments: MTVAF_EXTRA_NVCC_FLAGS="-DX=1" MTVAF_LIB_TAG=x python -m mtvaf_b200.build -> _C/libmtvaf_b200_x.so, picked up
# by mtvaf_b200.lib when MTVAF_LIB_TAG=x is set at import time
_TAG = os.environ.get("MTVAF_LIB_TAG", "")
if _TAG:
    LIB_PATH = os.path.join(OUT_DIR, "libmtvaf_b200_%s.so" % _TAG)
    OUT_DIR = os.path.join(OUT_DIR, "obj_" + _TAG)
FLAGS_EXTRA = os.environ.get("MTVAF_EXTRA_NVCC_FLAGS", "").split()
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(os.path.dirname(HERE), "include", "mtvaf_b200.h"), "rb").read())
    h.update(" ".join(FLAGS + FLAGS_EXTRA).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB_PATH
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, src[:-3] + ".o")
        cmd = [NVCC] + FLAGS + FLAGS_EXTRA + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(dig)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
