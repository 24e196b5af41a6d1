"""Data-parallel path on real GPUs: two ranks over NCCL, each with half of a global batch, must
end up with the same averaged gradients as one rank processing the whole batch (dropout off).  Needs 2 GPUs:
skipped on the single-GPU box of the round-end run, exercised with `gpurun --gpus 2`."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev):
    from oracle import mtvaf_oracle as O                 # config + init helpers only
    from oracle.make_golden import hf_config
    from mtvaf_b200 import synthetic as S
    from mtvaf_b200.modules import TVNetSAModel2, FeatureStub
    cfg = O.EncoderCfg.roberta_base(vocab_size=1000)
    params = S.init_params(cfg, seed=5, ln_jitter=0.05)
    args = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
                           beta=0.5, alpha=0.1, vao=True, noauxloss=False, resnet_root=None, compute_dtype="fp32")
    m = TVNetSAModel2(list(range(10)), None, args, config=hf_config(cfg), image_model=FeatureStub())
    m.load_state_dict(params, strict=False)
    return m.to(dev).eval()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from mtvaf_b200 import synthetic as S
        from mtvaf_b200.optim import GradSync
        full = S.make_batch(4, 32, vocab=1000, seed=9)
        # reference: the whole batch on this rank, no sync
        m = _build(dev)
        out, _, _ = m(**{k: v.to(dev) for k, v in full.items()})
        out.loss.backward()
        g_full = m.engine().flat.G.clone()
        # data parallel: half of the batch per rank, gradients averaged by GradSync
        m2 = _build(dev)
        sync = GradSync(m2.engine())
        half = {k: v[2 * rank:2 * rank + 2].to(dev) for k, v in full.items()}
        out2, _, _ = m2(**half)
        out2.loss.backward()
        sync.finish()
        torch.cuda.synchronize()
        g_dp = m2.engine().flat.G
        err = float((g_dp - g_full).abs().max() / (g_full.abs().max() + 1e-30))
        q.put((rank, err))
    except Exception as e:                               # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _worker_steps(rank, world, port, q):
    """K optimizer steps through the path bench.py times at N > 1: per-layer all-reduce hooks inside backward,
    `FlatAdamW.step(zero_grad=True, sync=...)` (tail all-reduce under the layer updates), eagerly and replayed from the
    whole-step CUDA graph -- against one rank stepping on the full batches."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from mtvaf_b200 import synthetic as S
        from mtvaf_b200.optim import GradSync, FlatAdamW
        from mtvaf_b200.graph import GraphedTrainStep
        K = 3
        fulls = [S.make_batch(4, 32, vocab=1000, seed=90 + i) for i in range(K)]
        halves = [{k: v[2 * rank:2 * rank + 2].to(dev) for k, v in b.items()} for b in fulls]

        def loss_of(out):
            return (out[0] if isinstance(out, tuple) else out).loss

        m = _build(dev)
        opt = FlatAdamW(m.engine(), lr=5e-5, warmup_steps=2, total_steps=10)
        for b in fulls:
            loss_of(m(**{k: v.to(dev) for k, v in b.items()})).backward()
            opt.step(zero_grad=True)
        W_full = m.engine().flat.W.clone()
        res = {}
        for mode in ("eager", "graph"):
            m2 = _build(dev)
            opt2 = FlatAdamW(m2.engine(), lr=5e-5, warmup_steps=2, total_steps=10)
            sync = GradSync(m2.engine(), optimizer=opt2)
            if mode == "eager":
                for b in halves:
                    loss_of(m2(**b)).backward()
                    opt2.step(zero_grad=True, sync=sync)
            else:
                g = GraphedTrainStep(m2, opt2, halves[0], grad_sync=sync, warmup=2)
                for b in halves:
                    g(b)
                torch.cuda.synchronize()
                g.close()
            torch.cuda.synchronize()
            W = m2.engine().flat.W
            f = m2.engine().flat
            worst = 0.0
            for n in f.names:
                o, k = f.offsets[n]
                nb = float(W_full[o:o + k].norm())
                if nb > 0:
                    worst = max(worst, float((W[o:o + k] - W_full[o:o + k]).norm()) / nb)
            chk = torch.stack([W.double().sum(), W.double().abs().sum()])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            res[mode] = (worst, bool(torch.equal(lo, hi)))
        q.put((rank, res))
    except Exception as e:                               # pragma: no cover
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_optimizer_steps_match_single_rank_and_ranks_stay_equal():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_steps, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, r in res:
        assert isinstance(r, dict), (rank, r)
        for mode, (worst, equal) in r.items():
            assert equal, (rank, mode, "ranks diverged")
            assert worst < 1e-4, (rank, mode, worst)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_match_single_rank_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, err in res:
        assert isinstance(err, float), (rank, err)
        assert err < 2e-4, (rank, err)
