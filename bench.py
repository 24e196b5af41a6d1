#!/usr/bin/env python
"""Headline benchmark: MTVAF RoBERTa-base bf16 TRAINING throughput on synthetic Twitter2017-shaped batches
(BASELINE.json configs[1]) -- one step = forward + backward + gradient sync + AdamW through the drop-in
TVNetSAModel2 (fusion P=16 + vao + probe + CRF), every op a kernel of mtvaf_b200.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                   (the oracle restatement of the reference on host cores)

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

L_TEXT = 128
N_AUX = 3
ALPHA, BETA = 0.1, 0.5
LR = 5e-5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mtvaf_b200", choices=["mtvaf_b200", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=int(os.environ.get("MTVAF_BENCH_BATCH", 512)),
                    help="samples per GPU and step (weak scaling); SURVEY.md 8(d) sweeps {16, 64, 256, 512}")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true",
                    help="issue every launch from Python each step instead of replaying the whole-step CUDA graph")
    ap.add_argument("--cpu-sample-batch", type=int, default=16)
    ap.add_argument("--deadline-s", type=float, default=float(os.environ.get("MTVAF_BENCH_DEADLINE_S", 1200)),
                    help="hard wall-clock limit: a hung collective must not outlive the round (exit code 3, no JSON line)")
    return ap.parse_args()


def flops_per_sample_step(L=L_TEXT, P=16, H=768, n=12, n_img=4, vao=True):
    """training step = 3 x forward algorithmic FLOPs."""
    Lk = P + L
    enc = n * L * (24 * H * H + 4 * Lk * H)
    fusion = n_img * 4 * 2 * (3840 * 800 + 800 * 8 * H) + n * n_img * 2 * 8 * H * 4 + (n_img * 2 * 8 * H * 2089 if vao else 0)
    probe = 2 * L * H * (H // 2)
    heads = 2 * L * H * 11
    return 3 * (enc + fusion + probe + heads)


def model_args(dtype):
    return SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
                           beta=BETA, alpha=ALPHA, vao=True, noauxloss=False, resnet_root=None, compute_dtype=dtype,
                           n_gpu=1)


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
def cpu_reference_run(batch_size: int, steps: int, warmup: int):
    """The reference's CPU path = the oracle restatement (oracle/mtvaf_oracle.py; the reference itself is
    Python and does not travel to the GPU box) on all host cores: fwd + bwd of TVNetSAModel2."""
    from oracle import mtvaf_oracle as O
    from mtvaf_b200 import synthetic as S
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.EncoderCfg.roberta_base()
    params = S.init_params(cfg, seed=1)
    params = {k: v.requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}
    batch = S.make_batch(batch_size, L_TEXT, shape="twitter2017", seed=2024)
    times = []
    for it in range(warmup + steps):
        for p in params.values():
            p.grad = None
        t0 = time.perf_counter()
        o = O.tvnet2_forward(params, cfg, batch, alpha=ALPHA, beta=BETA)
        o["loss"].backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return batch_size / med, med, torch.get_num_threads()


def run_reference(args, rank, world, emit):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1
    v, sec, cores = cpu_reference_run(args.cpu_sample_batch, steps, warm)
    line = {"impl": "reference", "metric": "train samples/s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MTVAF roberta-base TVNetSAModel2 training step (fusion P=16 + vao + probe + CRF), "
                                   "Twitter2017-shaped synthetic, L=128; CPU sample batch %d" % args.cpu_sample_batch},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": "%d fwd+bwd steps of batch %d, L=128 (oracle restatement of the reference, "
                                       "fp32, torch CPU eager)" % (steps, args.cpu_sample_batch)},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
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
        run_reference(args, rank, world, emit)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the mtvaf_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    layer_ctas = int(os.environ.get("MTVAF_NCCL_CTAS", 4))
    tail_ctas = int(os.environ.get("MTVAF_TAIL_CTAS", 16))
    reserve = int(os.environ.get("MTVAF_SM_RESERVE", 2 * layer_ctas))
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

    from transformers import RobertaConfig               # the product arm never imports oracle/
    from mtvaf_b200 import synthetic as S, ops
    from mtvaf_b200.modules import TVNetSAModel2, FeatureStub
    from mtvaf_b200.optim import FlatAdamW, GradSync

    B = args.per_gpu_batch
    # roberta-base (models/bert_model.py:425-429 loads it by name; no network here: random init of that architecture)
    hf_cfg = RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                           intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                           pad_token_id=1, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    torch.manual_seed(1234)                              # identical init on every rank
    model = TVNetSAModel2(list(range(10)), None, model_args(args.dtype), config=hf_cfg,
                          image_model=FeatureStub()).to(dev)
    model.train()
    eng = model.engine()
    eng.base_seed = 0x5EED + rank                        # different dropout streams per rank
    total_steps = args.warmup + 2 * args.steps + 8
    opt = FlatAdamW(eng, lr=LR, warmup_steps=max(1, total_steps // 100), total_steps=total_steps * 50)
    sync = GradSync(eng, tail_group=tail_group, optimizer=opt, reserve_sms=reserve,
                    tail_reserve_sms=tail_ctas if tail_group is not None else 0) if world > 1 else None
    dp_debug = os.environ.get("MTVAF_DP_DEBUG", "")          # diagnostics only (never set by the driver)
    if dp_debug == "nosync":
        eng.layer_grad_hook = None
        sync = None                                          # N independent replicas: isolates clock / power effects
    elif dp_debug == "tailonly" and sync is not None:
        eng.layer_grad_hook = None                           # everything reduced after backward: exposed comm time

    # distinct synthetic batches per rank (DistributedSampler-style disjoint shards), pinned on the host
    n_host = 4
    host = []
    for i in range(n_host):
        b = S.make_batch(B, L_TEXT, shape="twitter2017", seed=2024 + 100 * rank + i)
        host.append({k: v.pin_memory() for k, v in b.items()})
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    gemm_events = []
    ops.GEMM_EVENT_SINK = None

    def step(batch):
        out, prob, img = model(**batch)
        out.loss.backward()
        # AdamW + gradient clear in one pass over the flat buffers; data parallel: the tail all-reduce is launched
        # first and the encoder layers (reduced during backward) are updated beneath it
        opt.step(zero_grad=True, sync=sync)
        return out.loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager), then capture the whole step into one CUDA graph and warm the replay path
    graphed = None
    if args.no_graph:
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
        loss = run(resident[i % n_host])
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
    final_loss = float(loss.detach())

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
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    # DRAM bytes per GEMM launch: taken from the committed ncu capture of the SAME eager step (never measured under a
    # profiler here); null when the capture is for another batch size
    traffic, traffic_src = None, None
    try:
        cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles"))
                       if "gemm_traffic" in f and f.endswith("_b%d.json" % B))
        if cands and args.dtype == "bf16":
            tj = json.load(open(os.path.join(ROOT, "profiles", cands[-1])))
            traffic, traffic_src = tj["mean_dram_bytes_per_launch"], "profiles/" + cands[-1]
    except Exception:
        pass
    gemm_ms_per_step = g_ms / n_prof

    # ---- timed region 2: end to end through the public API with HOST buffers (pinned) -> `e2e`
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    if graphed is not None:
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
            l = step(cur)
            loss_host.copy_(l.detach().reshape(()), non_blocking=True)   # D2H read of the step's result
        e3.record()
        barrier()
    t = torch.tensor([e2.elapsed_time(e3)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) / 1e3)

    if rank == 0:
        line = {
            "metric": "train samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "MTVAF roberta-base TVNetSAModel2 training step (fwd+bwd+grad sync+AdamW; fusion "
                                   "P=16 + vao + probe + CRF), Twitter2017-shaped synthetic, L=128 (BASELINE.json "
                                   "configs[1])",
                       "per_gpu_batch": B, "global_batch": B * world, "seq_len": L_TEXT, "prefix_rows": 16,
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
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tc_kernel (tcgen05)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": gemm_alg_bytes / max(1, len(gemm_events)),
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback",
                         "gemm_launches_per_step": len(gemm_events) // n_prof,
                         "gemm_ms_per_step": gemm_ms_per_step,
                         "gemm_share_of_step": gemm_ms_per_step / ms_per_step if ms_per_step else None,
                         "eager_ms_per_step": eager_ms,
                         "step_model_tflops": flops_per_sample_step() * B / (ms_per_step * 1e-3) / 1e12},
        }
        if not args.no_cpu_baseline and world == 1:
            v, sec, cores = cpu_reference_run(args.cpu_sample_batch, 3, 1)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": "3 fwd+bwd steps of batch %d, L=128 (oracle restatement, fp32 torch "
                                              "CPU eager; no optimizer step)" % args.cpu_sample_batch}
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
