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
