"""BASELINE.json configs[4]: fusion-layer + psdProbe microbenchmark sweep on one B200 -- text length L in {64, 128, 256, 512}
x object regions (prefix rows) P in {10, 16, 36, 64, 100}, bf16, T = B * L = 32768 tokens per cell.

Per cell (CUDA events, L2 flushed between iterations, median of `iters`):
  attn_fwd / attn_bwd   prefix ("fusion") attention kernels alone (mtvaf_attention_fwd / _bwd_ex), with the kernel
                        path they dispatch to (tcgen05 pipelined / tcgen05 generic / SIMT) implied by the shape
  layer_fwd / layer_fb  ONE encoder layer through the drop-in RobertaModel with `past_key_values` (forward, forward+backward)
  probe_one             OneWordPSDProbe: x [T,768] @ proj [768,384] with the squared-norm epilogue (layers 4 / 7 share a shape)
  probe_two             TwoWordPSDProbe: pairwise squared distances [B, L, L] from the projected tokens
Once per run: the visual-prompt stack (get_visual_prompt: pyramid MLP + ANP heads + gates) at B = 256, 1 + 3 images.

  python tools/bench_sweep.py [--iters 10] [--out gpurun_out/sweep.json]

Not part of the test suite or of bench.py; results go under profiles/ by hand."""
import argparse
import json
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mtvaf_b200 import ops, lib as Lb

DEV = "cuda"
NH, D, H = 12, 64, 768
TOKENS = 32768


def timeit(fn, iters, flush):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)       # us
    ts.sort()
    return ts[len(ts) // 2]


def cell(Lq, P, iters, flush):
    B = max(1, TOKENS // Lq)
    T = B * Lq
    g = torch.Generator(device=DEV).manual_seed(Lq * 1000 + P)
    r = {"L": Lq, "P": P, "B": B}

    def rn(*s, scale=1.0):
        return (torch.randn(*s, device=DEV, generator=g) * scale).bfloat16()

    qkv = rn(T, 3 * H)
    kp, vp = rn(B, NH, P, D), rn(B, NH, P, D)
    lens = torch.randint(max(1, Lq // 4), Lq + 1, (B,), device=DEV, generator=g)
    mask = (torch.arange(Lq, device=DEV).unsqueeze(0) < lens.unsqueeze(1)).long()
    dctx = rn(T, H)
    dkp = torch.zeros(B, NH, P, D, device=DEV)
    dvp = torch.zeros(B, NH, P, D, device=DEV)
    dbias = torch.zeros(3 * H, device=DEV)
    st = {}

    def f():
        st["ctx"], st["lse"], _ = ops.attention_fwd(qkv, kp, vp, mask, B, Lq, NH, D, p_drop=0.1, seed=1)

    def b():
        ops.attention_bwd(dctx, qkv, kp, vp, mask, st["ctx"], st["lse"], B, Lq, NH, D, dkp=dkp, dvp=dvp, p_drop=0.1,
                          seed=1, d_bias=dbias)

    fl_f = 4.0 * B * NH * Lq * (P + Lq) * D
    try:
        us = timeit(f, iters, flush)
        r["attn_fwd_us"], r["attn_fwd_tflops"] = us, fl_f / us / 1e6
        us = timeit(b, iters, flush)
        r["attn_bwd_us"], r["attn_bwd_tflops"] = us, 2.5 * fl_f / us / 1e6
    except Exception as e:                                   # a shape a kernel rejects is a result, not a crash
        r["attn_error"] = repr(e)

    # one encoder layer with the prefix, through the drop-in module
    try:
        from transformers import RobertaConfig
        from mtvaf_b200.modules import RobertaModel
        cfg = RobertaConfig(vocab_size=4096, hidden_size=H, num_hidden_layers=1, num_attention_heads=NH,
                            intermediate_size=4 * H, max_position_embeddings=Lq + 2, type_vocab_size=1,
                            layer_norm_eps=1e-5, pad_token_id=1, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
        torch.manual_seed(0)
        m = RobertaModel.from_config(cfg, compute_dtype="bf16").to(DEV).train()
        ids = torch.randint(3, 4096, (B, Lq), device=DEV, generator=g) * mask
        full = torch.cat([torch.ones(B, P, device=DEV), mask.float()], 1)
        pkv = [(kp.float().requires_grad_(), vp.float().requires_grad_())]
        w = rn(B, Lq, H)

        def lf():
            with torch.no_grad():
                m(input_ids=ids, attention_mask=full, past_key_values=pkv, output_hidden_states=True, return_dict=True)

        def lfb():
            out = m(input_ids=ids, attention_mask=full, past_key_values=pkv, output_hidden_states=True, return_dict=True)
            (out["last_hidden_state"] * w).sum().backward()

        r["layer_fwd_us"] = timeit(lf, iters, flush)
        r["layer_fb_us"] = timeit(lfb, iters, flush)
        lk = P + Lq
        fl_layer = T * (24.0 * H * H + 4.0 * lk * H)
        r["layer_fwd_tflops"] = fl_layer / r["layer_fwd_us"] / 1e6
        r["layer_fb_tflops"] = 3.0 * fl_layer / r["layer_fb_us"] / 1e6
        del m
    except Exception as e:
        r["layer_error"] = repr(e)

    # probes (probes/probe.py:25-79): rank 384
    try:
        x = rn(T, H)
        proj = rn(H, 384, scale=0.05)
        norms = torch.zeros(T, device=DEV)
        Tm = torch.empty(T, 384, device=DEV, dtype=torch.bfloat16)

        def p1():
            ops.gemm(x, proj, b_mn=True, M=T, N=384, K=H, mode=Lb.EPI_SQNORM, rowvec=norms, out=Tm)

        r["probe_one_us"] = timeit(p1, iters, flush)
        Tf = Tm.float()

        def p2():
            ops.pairwise_sqdist(Tf, B, Lq, 384)

        r["probe_two_us"] = timeit(p2, iters, flush)
    except Exception as e:
        r["probe_error"] = repr(e)
    return r


def fusion(iters, flush, n_img=4):
    """get_visual_prompt at the reference shape (1 image + 3 aux crops) or with the full image only (n_img = 1):
    pyramid features [B, 3840, 2, 2], forward and forward+backward."""
    from transformers import RobertaConfig
    from mtvaf_b200.modules import TVNetSAModel2, FeatureStub
    B = 256
    args = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
                           beta=0.5, alpha=0.1, vao=True, noauxloss=False, resnet_root=None, compute_dtype="bf16", n_gpu=1,
                           probe_ckpt="")
    cfg = RobertaConfig(vocab_size=4096, hidden_size=H, num_hidden_layers=12, num_attention_heads=NH,
                        intermediate_size=4 * H, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                        pad_token_id=1)
    torch.manual_seed(0)
    m = TVNetSAModel2(list(range(10)), None, args, config=cfg, image_model=FeatureStub()).to(DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(5)
    images = torch.randn(B, 3840, 2, 2, device=DEV, generator=g).abs()
    aux = torch.randn(B, 3, 3840, 2, 2, device=DEV, generator=g).abs() if n_img > 1 else None
    label = torch.softmax(torch.randn(B, 2089, device=DEV, generator=g), -1)

    def f():
        with torch.no_grad():
            m.get_visual_prompt(images, aux, label)

    def fb():
        kv, l0, laux = m.get_visual_prompt(images, aux, label)
        (kv.float().sum() * 1e-3 + l0 + sum(laux)).backward()

    fl = B * n_img * (4 * 2.0 * (3840 * 800 + 800 * 8 * H) + 12 * 2.0 * 8 * H * 4 + 2.0 * 8 * H * 2089)
    r = {"B": B, "n_img": n_img, "visual_prompt_fwd_us": timeit(f, iters, flush)}
    r["visual_prompt_fwd_tflops"] = fl / r["visual_prompt_fwd_us"] / 1e6
    m.train()
    r["visual_prompt_fb_us"] = timeit(fb, iters, flush)
    r["visual_prompt_fb_tflops"] = 3 * fl / r["visual_prompt_fb_us"] / 1e6
    return r


def to_markdown(res):
    rows = ["| L | P | B | attn fwd | attn bwd | attn fwd TF/s | attn bwd TF/s | 1 layer fwd | 1 layer fwd+bwd | layer f+b TF/s | "
            "us per 1k tokens (layer f+b) | OneWord probe (layers 4 / 7: same shape) | TwoWord probe |",
            "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for c in res["cells"]:
        g = lambda k, f="%.0f": (f % c[k]) if k in c else "-"
        per1k = ("%.1f" % (c["layer_fb_us"] / (c["B"] * c["L"] / 1000.0))) if "layer_fb_us" in c else "-"
        rows.append("| %d | %d | %d | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
            c["L"], c["P"], c["B"], g("attn_fwd_us"), g("attn_bwd_us"), g("attn_fwd_tflops"), g("attn_bwd_tflops"),
            g("layer_fwd_us"), g("layer_fb_us"), g("layer_fb_tflops"), per1k, g("probe_one_us"), g("probe_two_us")))
    out = ["# configs[4] sweep -- `python tools/bench_sweep.py`, T = %d tokens per cell, bf16, us per call (median, L2 flushed)"
           % res["tokens_per_cell"], ""] + rows + [""]
    for f in res.get("fusion", []):
        out.append("Visual-prompt stack (`get_visual_prompt`, B=%d, n_img=%d): forward %.0f us (%.0f TF/s), forward+backward "
                   "%.0f us (%.0f TF/s)." % (f["B"], f["n_img"], f["visual_prompt_fwd_us"], f["visual_prompt_fwd_tflops"],
                                             f["visual_prompt_fb_us"], f["visual_prompt_fb_tflops"]))
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--L", type=int, nargs="*", default=[64, 128, 256, 512])
    ap.add_argument("--P", type=int, nargs="*", default=[10, 16, 36, 64, 100])
    a = ap.parse_args()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    res = {"cells": [], "tokens_per_cell": TOKENS, "dtype": "bf16", "l2": "256 MB buffer written between iterations"}
    for Lq in a.L:
        for P in a.P:
            c = cell(Lq, P, a.iters, flush)
            print(json.dumps(c), flush=True)
            res["cells"].append(c)
    res["fusion"] = []
    for n_img in (1, 4):
        try:
            res["fusion"].append(fusion(a.iters, flush, n_img))
            print(json.dumps(res["fusion"][-1]), flush=True)
        except Exception as e:
            res["fusion_error_%d" % n_img] = repr(e)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    open(os.path.splitext(a.out)[0] + ".md", "w").write(to_markdown(res))


if __name__ == "__main__":
    main()
