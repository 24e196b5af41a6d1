"""Prefix-attention kernels alone at one shape: `python tools/bench_attention.py [B L P]` -> us per launch (CUDA events
around 20 launches on the launch stream, inputs (3 x B x L x 768 bf16) larger than L2).  The bench shape is `512 128 16`;
the default `512 64 16` is the short-text shape that runs two (batch, head) items per tile."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mtvaf_b200 import ops

B, L, P = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (512, 64, 16)))
NH, D = 12, 64
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
bf = torch.bfloat16
qkv = torch.randn(B * L, 3 * NH * D, device=dev, generator=g).to(bf)
kp = torch.randn(B, NH, P, D, device=dev, generator=g).to(bf) if P else None
vp = torch.randn(B, NH, P, D, device=dev, generator=g).to(bf) if P else None
mask = torch.ones(B, L, dtype=torch.long, device=dev)
dctx = torch.randn(B * L, NH * D, device=dev, generator=g).to(bf)
dkp = torch.zeros(B, NH, P, D, device=dev) if P else None
dvp = torch.zeros(B, NH, P, D, device=dev) if P else None


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


st = {}


def fwd():
    st["ctx"], st["lse"], _ = ops.attention_fwd(qkv, kp, vp, mask, B, L, NH, D, p_drop=0.1, seed=1)


def bwd():
    ops.attention_bwd(dctx, qkv, kp, vp, mask, st["ctx"], st["lse"], B, L, NH, D, dkp, dvp, p_drop=0.1, seed=1)


print("B=%d L=%d P=%d  attn_fwd %.1f us  attn_bwd %.1f us" % (B, L, P, timeit(fwd), timeit(bwd)))
