#!/usr/bin/env python
"""Headline benchmark: MTVAF RoBERTa-base bf16 TRAINING throughput on synthetic Twitter2017-shaped batches
(BASELINE.json configs[1]) -- one step = forward + backward + gradient sync + AdamW through the drop-in
TVNetSAModel2 (fusion P=16 + vao + probe + CRF), every op a kernel of mtvaf_b200.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                   (the UNMODIFIED reference on the host cores, from oracle/_ref)
  python bench.py --config large_l256 | infer_bs512      (BASELINE.json configs[2] / configs[3])
  python bench.py --global-batch G --gpus N              (strong scaling: G samples per step over all ranks)

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_AUX = 3
ALPHA, BETA = 0.1, 0.5
LR = 5e-5

# BASELINE.json configs -> workloads.  `train` is the configuration the metric is quoted on (configs[1]).
WORKLOADS = {
    "train": dict(model="roberta-base", H=768, layers=12, heads=12, inter=3072, L=128, batch=512, mode="train",
                  prefix=True, probe=True, vao=True, shape="twitter2017", cpu_batch=16, metric="train samples/s",
                  text="MTVAF roberta-base TVNetSAModel2 training step (fwd+bwd+grad sync+AdamW; fusion P=16 + vao + "
                       "probe + CRF), Twitter2017-shaped synthetic, L=128 (BASELINE.json configs[1])"),
    # roberta-large + prefix/probe has NO reference behaviour (12/768 hard-coded in models/bert_model.py:229,455,544;
    # SURVEY.md section 0): the config the reference can run is text + CRF head, which is what both arms run here
    "large_l256": dict(model="roberta-large", H=1024, layers=24, heads=16, inter=4096, L=256, batch=128, mode="train",
                       prefix=False, probe=False, vao=False, shape="longaux", cpu_batch=2, metric="train samples/s",
                       text="MTVAF roberta-large TVNetSAModel2 training step (fwd+bwd+grad sync+AdamW; text with long "
                            "auxiliary context + CRF head, use_prefix/use_probe off as the reference requires at "
                            "large), L=256 (BASELINE.json configs[2])"),
    "infer_bs512": dict(model="roberta-base", H=768, layers=12, heads=12, inter=3072, L=128, batch=512, mode="infer",
                        prefix=True, probe=True, vao=True, shape="twitter2017", cpu_batch=16, metric="infer samples/s",
                        text="MTVAF roberta-base TVNetSAModel2 eval forward (model.eval(), no_grad, labels passed as the "
                             "reference's evaluate()/test() do; fusion P=16 + probe + CRF NLL + Viterbi decode), "
                             "bs=512, L=128 (BASELINE.json configs[3])"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mtvaf_b200", choices=["mtvaf_b200", "reference"])
    ap.add_argument("--config", default="train", choices=sorted(WORKLOADS))
    ap.add_argument("--per-gpu-batch", type=int, default=int(os.environ.get("MTVAF_BENCH_BATCH", 0)),
                    help="samples per GPU and step (weak scaling); default: the workload's (512 for `train`; "
                         "SURVEY.md 8(d) sweeps {16, 64, 256, 512})")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="STRONG scaling: total samples per step, split evenly over the ranks")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true",
                    help="skip timing the unmodified reference modules in PyTorch eager on this GPU (N=1, `train` only)")
    ap.add_argument("--no-graph", action="store_true",
                    help="issue every launch from Python each step instead of replaying the whole-step CUDA graph")
    ap.add_argument("--cpu-sample-batch", type=int, default=0)
    ap.add_argument("--deadline-s", type=float, default=float(os.environ.get("MTVAF_BENCH_DEADLINE_S", 1500)),
                    help="hard wall-clock limit: a hung collective must not outlive the round (exit code 3, no JSON line)")
    return ap.parse_args()


def flops_per_sample(w, train=True, P=16, n_img=4):
    """forward algorithmic FLOPs per sample; a training step = 3 x forward."""
    H, n, L, I = w["H"], w["layers"], w["L"], w["inter"]
    Lk = (P if w["prefix"] else 0) + L
    enc = n * L * (8 * H * H + 4 * H * I + 4 * Lk * H)           # QKV 6H^2 + O 2H^2 + FFN 4HI | QK^T + PV
    fusion = 0
    if w["prefix"]:
        fusion = n_img * 4 * 2 * (3840 * 800 + 800 * 8 * H) + n * n_img * 2 * 8 * H * 4
        if w["vao"]:
            fusion += n_img * 2 * 8 * H * 2089
    probe = 2 * L * H * (H // 2) if w["probe"] else 0
    heads = 2 * L * H * 11
    return (3 if train else 1) * (enc + fusion + probe + heads)


def model_args(w, dtype):
    return SimpleNamespace(bert_name=w["model"], prefix_dim=768, prefix_len=4, use_prefix=w["prefix"],
                           use_probe=w["probe"], beta=BETA, alpha=ALPHA, vao=w["vao"], noauxloss=False,
                           resnet_root=None, compute_dtype=dtype, n_gpu=1, probe_ckpt="")


def hf_roberta_config(w):
    from transformers import RobertaConfig
    # models/bert_model.py:425-429 loads the backbone by name; no network here: random init of that architecture
    return RobertaConfig(vocab_size=50265, hidden_size=w["H"], num_hidden_layers=w["layers"],
                         num_attention_heads=w["heads"], intermediate_size=w["inter"], max_position_embeddings=514,
                         type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1, hidden_dropout_prob=0.1,
                         attention_probs_dropout_prob=0.1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# baselines: the reference's own implementation (checker code under oracle/ -- never on the product arm)
# ------------------------------------------------------------------------------------------------
def _reference_model(w, device):
    """The UNMODIFIED reference TVNetSAModel2 (oracle/_ref copy or /root/reference, through oracle/ref_shim.py), random
    init of the workload's architecture.  Returns (model, kind) or (None, why)."""
    try:
        from oracle import ref_shim
        if not ref_shim.reference_available():
            return None, "reference tree not staged (oracle/_ref missing)"
        rargs = ref_shim.make_args(bert_name=w["model"], use_prefix=w["prefix"], use_probe=w["probe"], vao=w["vao"],
                                   device=device)
        model = ref_shim.build_reference_tvnet2(hf_roberta_config(w), rargs, list(range(10)), seed=1)
        return model.to(device), "reference"
    except Exception as e:                                       # noqa: BLE001
        return None, "%s: %s" % (type(e).__name__, e)


def _ref_kwargs(w, batch, with_labels=True):
    kw = dict(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
              token_type_ids=batch["token_type_ids"])
    if with_labels:
        kw["labels"] = batch["labels"]
    if w["prefix"]:
        kw.update(images=batch["images"], aux_imgs=batch["aux_imgs"], imagelabel=batch["imagelabel"])
    return kw


def cpu_reference_run(w, batch_size: int, steps: int, warmup: int):
    """The reference's CPU path on all host cores: the unmodified reference modules when staged (kind "reference"),
    else the oracle restatement (kind "port").  train: fwd + bwd of TVNetSAModel2 (train mode, dropout on -- as the
    reference trains); infer: eval forward under no_grad.  Returns (samples/s, s/step, threads, kind, note)."""
    from mtvaf_b200 import synthetic as S
    torch.set_num_threads(os.cpu_count() or 1)
    batch = S.make_batch(batch_size, w["L"], shape=w["shape"], seed=2024, with_images=w["prefix"])
    train = w["mode"] == "train"
    model, kind = _reference_model(w, torch.device("cpu"))
    note = "unmodified reference modules (oracle/_ref), fp32 torch CPU eager; torchcrf absent -> CRF = oracle restatement"
    if model is None:
        from oracle import mtvaf_oracle as O
        note = "oracle restatement (reference unavailable: %s), fp32 torch CPU eager" % kind
        kind = "port"
        cfg = O.EncoderCfg(kind="roberta", hidden_size=w["H"], num_hidden_layers=w["layers"],
                           num_attention_heads=w["heads"], intermediate_size=w["inter"])
        params = S.init_params(cfg, seed=1, with_fusion=w["prefix"])
        params = {k: v.requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}

        def one():
            for p in params.values():
                p.grad = None
            with torch.set_grad_enabled(train):
                o = O.tvnet2_forward(params, cfg, batch, use_prefix=w["prefix"], use_probe=w["probe"], vao=w["vao"], alpha=ALPHA, beta=BETA)
                if train:
                    o["loss"].backward()
    else:
        model.train(train)

        def one():
            model.zero_grad(set_to_none=True)
            with torch.set_grad_enabled(train):
                out = model(**_ref_kwargs(w, batch))          # eval passes labels too (modules/train.py:696-857)
                if train:
                    (out[0] if isinstance(out, tuple) else out).loss.backward()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return batch_size / med, med, torch.get_num_threads(), kind, note


def gpu_eager_reference_run(w, dev, B, steps=3, warmup=2):
    """"the real kernel to beat" -- the unmodified reference modules in PyTorch eager
    (cuBLAS / ATen) on THIS GPU: fp32 as shipped, bf16 via torch.autocast.  Training step = fwd + bwd + torch AdamW
    with the reference's groups.  `as_shipped` keeps use_probe=True, whose ConstructLabelGaget is a Python double loop
    over 0-d CUDA tensors (probes/constructLabel.py:14-28, one device sync per comparison): timed at a smaller batch,
    its cost is per sample.  The other two arms switch the probe off so that the incumbent's KERNELS are what is timed."""
    from mtvaf_b200 import synthetic as S
    res = {}

    def timed(model, batch, autocast):
        params = [p for n, p in model.named_parameters() if "image_model" not in n]
        opt = torch.optim.AdamW(params, lr=LR)

        def one():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                out = model(**_ref_kwargs(w, batch))
                loss = (out[0] if isinstance(out, tuple) else out).loss
            loss.backward()
            opt.step()
            opt.zero_grad()
        for _ in range(warmup):
            one()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for name, probe, autocast, b in (("fp32", False, False, B), ("bf16_autocast", False, True, B),
                                     ("fp32_as_shipped_with_probe", True, False, min(B, 32))):
        w2 = dict(w, probe=probe and w["probe"])
        if name.endswith("with_probe") and not w["probe"]:
            continue
        try:
            model, kind = _reference_model(w2, dev)
            if model is None:
                res[name] = {"unavailable": kind}
                continue
            model.train()
            batch = {k: v.to(dev) for k, v in S.make_batch(b, w["L"], shape=w["shape"], seed=2024,
                                                           with_images=w["prefix"]).items()}
            ms = timed(model, batch, autocast)
            res[name] = {"value": b / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms, "batch": b,
                         "use_probe": bool(w2["probe"])}
            del model, batch
            torch.cuda.empty_cache()
        except Exception as e:                                   # noqa: BLE001
            res[name] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
            torch.cuda.empty_cache()
    res["note"] = ("unmodified reference modules (oracle/_ref) in PyTorch eager on this GPU, training step = fwd + bwd + "
                   "torch.optim.AdamW; CRF = oracle restatement (torchcrf absent); dropout on")
    return res


def run_reference(args, w, rank, world, emit):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1
    cb = args.cpu_sample_batch or w["cpu_batch"]
    v, sec, cores, kind, note = cpu_reference_run(w, cb, steps, warm)
    line = {"impl": "reference", "metric": w["metric"], "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["text"] + "; CPU sample batch %d" % cb, "name": args.config},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
                             "sample": "%d steps of batch %d, L=%d: %s" % (steps, cb, w["L"], note)},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------
GEMM_NAMES = {}


def name_gemm(w, B, key):
    """Label a tcgen05 GEMM launch by its role on the path (shape-keyed; key = (M, N, K, a_mn, b_mn, mode))."""
    M, N, K, a_mn, b_mn, mode = key
    H, I, T = w["H"], w["inter"], B * w["L"]
    rows4, rows = 16 * B, 4 * B
    role = None
    if (a_mn, b_mn) == (0, 0):
        role = {(T, 3 * H, H): "qkv_fwd", (T, H, H): "attn_out_fwd_resid", (T, I, H): "ffn1_fwd_gelu(+grad)",
                (T, H, I): "ffn2_fwd_resid", (rows4, 800, 3840): "fusion_mlp1_fwd_tanh",
                (rows4, 8 * H, 800): "fusion_mlp2_fwd", (B, 2089, 8 * H): "anp_head_fwd",
                (rows, 4 * w["layers"], 8 * H): "fusion_gate_logits_fwd", (T, 11, H): "tag_head_fwd",
                (T, H, H // 2): "probe_dgrad"}.get((M, N, K))
    elif (a_mn, b_mn) == (0, 1):
        role = {(T, H, 3 * H): "qkv_dgrad_resid", (T, H, H): "attn_out_dgrad", (T, H, I): "ffn1_dgrad_resid",
                (T, I, H): "ffn2_dgrad_mulgelugrad_colsum", (T, H // 2, H): "probe_fwd",
                (rows4, 800, 8 * H): "fusion_mlp2_dgrad_dtanh", (B, 8 * H, 2089): "anp_head_dgrad",
                (T, H, 11): "tag_head_dgrad"}.get((M, N, K))
    elif (a_mn, b_mn) == (1, 1):
        role = {(3 * H, H, T): "qkv_wgrad", (H, H, T): "attn_out_wgrad", (I, H, T): "ffn1_wgrad",
                (H, I, T): "ffn2_wgrad", (8 * H, 800, rows4): "fusion_mlp2_wgrad", (800, 3840, rows4): "fusion_mlp1_wgrad",
                (2089, 8 * H, B): "anp_head_wgrad", (11, H, T): "tag_head_wgrad", (H, H // 2, T): "probe_wgrad"}.get((M, N, K))
    return role or "other_%dx%dx%d_%d%d_m%d" % key


def main():
    args = parse()
    w = WORKLOADS[args.config]
    if args.deadline_s > 0:
        def _deadline():
            sys.stderr.write("bench.py: exceeded --deadline-s %.0f s (hung collective or device?) -- aborting\n" % args.deadline_s)
            sys.stderr.flush()
            os._exit(3)
        _t = threading.Timer(args.deadline_s, _deadline)
        _t.daemon = True
        _t.start()
    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner) write there too, so everything
    # before the final print goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, w, rank, world, emit)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the mtvaf_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    layer_ctas = int(os.environ.get("MTVAF_NCCL_CTAS", 4))
    tail_ctas = int(os.environ.get("MTVAF_TAIL_CTAS", 16))
    # SMs the persistent kernels leave to the layer all-reduces: one per NCCL CTA (measured: 2 GPUs 42.69 -> 41.94 ms per step,
    # 8 GPUs 43.80 -> 43.46 ms against twice that; 2 CTAs instead of 4 is faster still at N = 2 but slower at N = 8)
    reserve = int(os.environ.get("MTVAF_SM_RESERVE", layer_ctas))
    tail_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # Two communicators.  The per-layer gradient all-reduces have the rest of backward to hide behind (340 MB per
        # ~15 ms): `layer_ctas` CTAs are plenty, and the persistent tcgen05 kernels size their grids around them
        # (GradSync reserve_sms).  The tail (embedding tables, fusion MLP: final only when backward ends) is exposed,
        # so it goes out on a wide communicator while AdamW updates the layers.
        def nccl_opts(max_ctas):
            try:
                o = dist.ProcessGroupNCCL.Options()
                o.config.max_ctas = max_ctas
                o.config.min_ctas = 1
                return o
            except Exception:
                return None
        o1 = nccl_opts(layer_ctas)
        if o1 is None:
            os.environ.setdefault("NCCL_MAX_CTAS", str(layer_ctas))
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("nccl", device_id=dev, pg_options=o1)
            try:
                tail_group = dist.new_group(pg_options=nccl_opts(tail_ctas))
            except Exception:
                tail_group = None

    from mtvaf_b200 import synthetic as S, ops           # the product arm never imports oracle/
    from mtvaf_b200.modules import TVNetSAModel2, FeatureStub
    from mtvaf_b200.optim import FlatAdamW, GradSync

    scaling = "weak"
    B = args.per_gpu_batch or w["batch"]
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch must be a multiple of the number of ranks")
        B = args.global_batch // world
        scaling = "strong"
    L_TEXT = w["L"]
    train = w["mode"] == "train"
    torch.manual_seed(1234)                              # identical init on every rank
    model = TVNetSAModel2(list(range(10)), None, model_args(w, args.dtype), config=hf_roberta_config(w),
                          image_model=FeatureStub() if w["prefix"] else None).to(dev)
    model.train(train)
    eng = model.engine()
    eng.base_seed = 0x5EED + rank                        # different dropout streams per rank
    opt = sync = None
    dp_debug = os.environ.get("MTVAF_DP_DEBUG", "")          # diagnostics only (never set by the driver)
    if train:
        total_steps = args.warmup + 2 * args.steps + 8
        # use_prefix: the reference's name-selected groups (modules/train.py:894-926); text-only: AdamW over every
        # parameter at args.lr (`bert_before_train`, :887-892)
        groups = None if w["prefix"] else [(lambda n: True, LR, 1e-2)]
        opt = FlatAdamW(eng, lr=LR, groups=groups, warmup_steps=max(1, total_steps // 100), total_steps=total_steps * 50)
        sync = GradSync(eng, tail_group=tail_group, optimizer=opt, reserve_sms=reserve,
                        tail_reserve_sms=tail_ctas if tail_group is not None else 0) if world > 1 else None
        if dp_debug == "nosync":
            eng.layer_grad_hook = None
            sync = None                                      # N independent replicas: isolates clock / power effects
        elif dp_debug == "tailonly" and sync is not None:
            eng.layer_grad_hook = None                       # everything reduced after backward: exposed comm time

    # distinct synthetic batches per rank (DistributedSampler-style disjoint shards), pinned on the host
    n_host = 4
    host = []
    for i in range(n_host):
        b = S.make_batch(B, L_TEXT, shape=w["shape"], seed=2024 + 100 * rank + i, with_images=w["prefix"])
        if w["prefix"] and args.dtype == "bf16":
            # visual features travel in the bf16 wire format of the offline front-end stage (mtvaf_b200/features.py):
            # the fusion GEMM consumes bf16 anyway, so this is bit-identical to feeding fp32 features in bf16 mode
            b["images"], b["aux_imgs"] = b["images"].to(torch.bfloat16), b["aux_imgs"].to(torch.bfloat16)
        host.append({k: v.pin_memory() for k, v in b.items()})
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    gemm_events = []
    ops.GEMM_EVENT_SINK = None

    def loss_of(out):
        return (out[0] if isinstance(out, tuple) else out).loss

    def train_step(batch):
        loss = loss_of(model(**batch))
        loss.backward()
        # AdamW + gradient clear in one pass over the flat buffers; data parallel: the tail all-reduce is launched
        # first and the encoder layers (reduced during backward) are updated beneath it
        opt.step(zero_grad=True, sync=sync)
        return loss

    def infer_step(batch):
        with torch.no_grad():
            out = model(**batch)
        tags = (out[0] if isinstance(out, tuple) else out).logits       # CRF decode, lazily materialised
        return tags.device_tags                                         # (best [B,L] int64, lens [B] int64) on device

    step = train_step if train else infer_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager), then capture the whole step into one CUDA graph and warm the replay path
    graphed = None
    if args.no_graph or not train:
        for i in range(args.warmup):
            step(resident[i % n_host])
        run = step
    else:
        from mtvaf_b200.graph import GraphedTrainStep
        graphed = GraphedTrainStep(model, opt, resident[0], grad_sync=sync, warmup=args.warmup)
        for i in range(args.warmup):
            graphed(resident[i % n_host])
        run = graphed
    barrier()

    # ---- timed region 1: inputs resident in HBM -> `value`
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        res = run(resident[i % n_host])
    e1.record()
    barrier()
    launches = ops.launch_count() if graphed is None else graphed.kernels_per_replay * args.steps
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)
    final_loss = float(res.detach()) if train else None

    # ---- dominant kernel: the tcgen05 GEMM, timed per launch with CUDA events on the launch stream in a short
    # EAGER pass (events cannot be read inside a graph replay).  achieved = algorithmic 2*M*N*K summed over the
    # launches / summed event durations of those launches
    n_prof = max(1, min(args.steps, 3))
    if graphed is not None:
        step(resident[0])                 # untimed: lets the caching allocator size its eager pools after the capture
    ops.GEMM_EVENT_SINK = gemm_events
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for i in range(n_prof):
        step(resident[i % n_host])
    p1.record()
    barrier()
    ops.GEMM_EVENT_SINK = None
    eager_ms = p0.elapsed_time(p1) / n_prof
    g_fl = sum(e[2] for e in gemm_events)
    gemm_alg_bytes = sum(e[3] for e in gemm_events)      # operands + results (+ fused aux / second output) once each
    g_ms = sum(e[0].elapsed_time(e[1]) for e in gemm_events)
    by_role = {}
    for e in gemm_events:
        role = name_gemm(w, B, e[4])
        r = by_role.setdefault(role, {"launches_per_step": 0, "ms_per_step": 0.0, "flop": 0.0, "MNK": list(e[4][:3])})
        r["launches_per_step"] += 1
        r["ms_per_step"] += e[0].elapsed_time(e[1])
        r["flop"] += e[2]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    burst = peaks.get("bf16_tflops", None)
    gemms = {}
    for role, r in sorted(by_role.items(), key=lambda kv: -kv[1]["ms_per_step"]):
        tf = r["flop"] / (r["ms_per_step"] * 1e-3) / 1e12 if r["ms_per_step"] > 0 else 0.0
        gemms[role] = {"MNK": r["MNK"], "launches_per_step": r["launches_per_step"] // n_prof,
                       "ms_per_step": round(r["ms_per_step"] / n_prof, 4), "tflops": round(tf, 1),
                       "frac_of_burst_peak": round(tf / burst, 3) if burst else None}
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    # DRAM bytes per GEMM launch: taken from the committed ncu capture of the SAME eager step (never measured under a
    # profiler here); null when the capture is for another batch size / workload
    traffic, traffic_src = None, None
    try:
        cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles"))
                       if "gemm_traffic" in f and f.endswith("_b%d.json" % B))
        if cands and args.dtype == "bf16" and args.config == "train":
            tj = json.load(open(os.path.join(ROOT, "profiles", cands[-1])))
            traffic, traffic_src = tj["mean_dram_bytes_per_launch"], "profiles/" + cands[-1]
    except Exception:
        pass
    gemm_ms_per_step = g_ms / n_prof

    # ---- timed region 2: end to end through the public API with HOST buffers (pinned) -> `e2e`
    d2h_bytes = 4
    if graphed is not None:
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        graphed.prefetch(host[0])
        for i in range(args.steps):
            l = graphed(None)                                 # consume the staged batch, replay the step
            if i + 1 < args.steps:
                graphed.prefetch(host[(i + 1) % n_host])      # H2D of the next batch behind this replay
            loss_host.copy_(l.reshape(()), non_blocking=True)         # D2H read of the step's result
        e3.record()
        barrier()
    else:
        copy_stream = torch.cuda.Stream()
        if train:
            result_host = torch.empty((), dtype=torch.float32).pin_memory()
        else:
            result_host = (torch.empty((B, L_TEXT), dtype=torch.int64).pin_memory(),
                           torch.empty((B,), dtype=torch.int64).pin_memory())
            d2h_bytes = 8 * B * L_TEXT + 8 * B

        def h2d(hb):
            with torch.cuda.stream(copy_stream):
                d = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return d, ev

        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        nxt = h2d(host[0])
        for i in range(args.steps):
            cur, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            for v in cur.values():                            # allocated on the copy stream, consumed on this one
                v.record_stream(torch.cuda.current_stream())
            if i + 1 < args.steps:
                nxt = h2d(host[(i + 1) % n_host])            # prefetch the next batch behind this step's compute
            r = step(cur)
            if train:
                result_host.copy_(r.detach().reshape(()), non_blocking=True)   # D2H read of the step's result
            else:
                result_host[0].copy_(r[0], non_blocking=True)                  # decoded tags + lengths
                result_host[1].copy_(r[1], non_blocking=True)
        e3.record()
        barrier()
    t = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) / 1e3)

    # ---- data parallel: the ranks must have stayed in lock-step through opt.step(sync=...) for every step above
    dp_check = None
    if world > 1 and train:
        f = eng.flat
        chk = torch.stack([f.W.double().sum(), f.W.double().abs().sum(), opt.m.double().abs().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dp_check = {"weights_equal": bool(torch.equal(lo[:2], hi[:2])), "adam_m_equal": bool(lo[2] == hi[2]),
                    "checksum": float(lo[0]), "max_rank_spread": float((hi - lo).abs().max()),
                    "steps_checked": int(opt.dyn[0]) if opt.dyn is not None else int(opt.t), "synced": dp_debug != "nosync"}

    if rank == 0:
        fl = flops_per_sample(w, train=train)
        line = {
            "metric": w["metric"], "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": w["text"], "name": args.config,
                       "per_gpu_batch": B, "global_batch": B * world, "seq_len": L_TEXT,
                       "prefix_rows": 16 if w["prefix"] else 0,
                       "parallelism": "dp%d" % world,
                       "dp": None if world == 1 else {"debug": dp_debug or None, "layer_allreduce_ctas": layer_ctas, "tail_allreduce_ctas": tail_ctas
                                                      if tail_group is not None else layer_ctas,
                                                      "sm_reserve_during_backward": reserve,
                                                      "reduced": "optimizer-owned ranges only"},
                       "l2": "inputs+activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                       "launch": "eager (one Python/ctypes call per kernel)" if graphed is None else
                                 "whole step (fwd+bwd+all-reduce+AdamW) replayed from one CUDA graph",
                       "final_loss": final_loss},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tc2_kernel (tcgen05 cta_group::2; gemm_bf16_tc_kernel "
                                                      "for M < 256)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": gemm_alg_bytes / max(1, len(gemm_events)),
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback",
                         "burst_peak": burst, "frac_of_burst": achieved / burst if burst else None,
                         "gemm_launches_per_step": len(gemm_events) // n_prof,
                         "gemm_ms_per_step": gemm_ms_per_step,
                         "gemm_share_of_step": gemm_ms_per_step / ms_per_step if ms_per_step else None,
                         "eager_ms_per_step": eager_ms,
                         "step_model_tflops": fl * B / (ms_per_step * 1e-3) / 1e12,
                         "gemms": gemms},
        }
        if dp_check is not None:
            line["dp_check"] = dp_check
        if world == 1 and args.config == "train" and not args.no_gpu_eager_baseline:
            if graphed is not None:
                graphed.close()
                graphed = None
            model = opt = None
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_reference_run(w, dev, B)
        if not args.no_cpu_baseline and world == 1:
            cb = args.cpu_sample_batch or w["cpu_batch"]
            v, sec, cores, kind, note = cpu_reference_run(w, cb, 2, 1)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
                                    "sample": "2 steps of batch %d, L=%d: %s" % (cb, L_TEXT, note)}
        emit(line)
    if world > 1:
        # tear down in dependency order: the captured graph holds NCCL kernels of this communicator
        if graphed is not None:
            graphed.close()
            graphed = None
        barrier()
        watchdog = threading.Timer(30.0, lambda: os._exit(0))   # never let a teardown hang outlive the result
        watchdog.daemon = True
        watchdog.start()
        dist.destroy_process_group()
        watchdog.cancel()


if __name__ == "__main__":
    main()
