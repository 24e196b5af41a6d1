"""Host-side logic of the data-parallel path on CPU: world_size-2 `gloo` process group.

GradSync replaces the reference's broken modules/parallel.py: every slice of the flat gradient buffer must be
all-reduced (averaged) exactly once per step, whether its layer hook fired during backward or not."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_engine(total, layer_ranges, rank):
    g = torch.Generator().manual_seed(100 + rank)
    G = torch.randn(total, generator=g)
    flat = SimpleNamespace(G=G, layer_ranges=layer_ranges, total=total)
    return SimpleNamespace(flat=flat, layer_grad_hook=None, tail_grad_hook=None)


def _worker(rank, world, port, fired, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mtvaf_b200.optim import GradSync
        layer_ranges = [(64, 200), (256, 400), (448, 640)]       # gaps = alignment padding
        total = 1024
        eng = _fake_engine(total, layer_ranges, rank)
        expect = sum(_fake_engine(total, layer_ranges, r).flat.G for r in range(world)) / world
        sync = GradSync(eng)
        assert eng.layer_grad_hook is not None                   # hook installed for world > 1
        for step in range(2):                                    # state must reset between steps
            if step == 1:
                eng.flat.G.copy_(_fake_engine(total, layer_ranges, rank).flat.G)
            for i in fired:                                      # backward order: last layer first
                eng.layer_grad_hook(i)
            sync.finish()
            covered = torch.zeros(total, dtype=torch.bool)
            covered[:64] = True
            for a, b in layer_ranges:
                covered[a:b] = True
            covered[640:] = True
            assert torch.allclose(eng.flat.G[covered], expect[covered], atol=1e-6), "step %d" % step
            assert not sync.works and not sync.tail_works and not sync.done_layers
        q.put((rank, "ok"))
    except Exception as e:                                       # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _worker_owned(rank, world, port, q):
    """GradSync(optimizer=...): only the ranges the optimizer updates are reduced; launch_tail / wait_layers /
    wait_tail (the split FlatAdamW.step(sync=...) uses) cover every owned element exactly once."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mtvaf_b200.optim import GradSync
        layer_ranges = [(0, 4096), (4096, 8192)]
        total = 32768
        # owned: both layers, a head range right after them, and the embedding tables at the end; the 8192+2048..20000
        # stretch (ANP heads / projectors / probe in the real model) is never stepped
        owned = [(0, 8192, 5e-5, 1e-2), (8192, 10240, 5e-2, 1e-2), (20000, 32768, 5e-5, 1e-2)]
        opt = SimpleNamespace(ranges=owned)
        eng = _fake_engine(total, layer_ranges, rank)
        mine = eng.flat.G.clone()
        expect = sum(_fake_engine(total, layer_ranges, r).flat.G for r in range(world)) / world
        sync = GradSync(eng, optimizer=opt)
        assert sync.tail_ranges() == [(8192, 10240), (20000, 32768)]
        eng.layer_grad_hook(1)
        sync.launch_tail()                                       # layer 0 never fired: must be picked up here
        sync.wait_layers()
        sync.wait_tail()
        m = torch.zeros(total, dtype=torch.bool)
        for a, b, _, _ in owned:
            m[a:b] = True
        assert torch.allclose(eng.flat.G[m], expect[m], atol=1e-6)
        assert torch.equal(eng.flat.G[~m], mine[~m])             # rank-local: untouched
        assert not sync.works and not sync.tail_works and not sync.done_layers
        q.put((rank, "ok"))
    except Exception as e:                                       # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _worker_embeddings(rank, world, port, q):
    """The embedding tables (packed last in the flat buffer) are reduced from `tail_grad_hook`, inside backward; the
    tail pass at the end must then skip them (no double averaging) and the flag must reset for the next step."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mtvaf_b200.optim import GradSync
        layer_ranges = [(0, 1024), (1024, 2048)]
        total = 8192
        names = ["bert.encoder.layer.0.w", "bert.encoder.layer.1.w", "fc.weight", "bert.embeddings.word_embeddings.weight",
                 "bert.embeddings.LayerNorm.weight"]
        offsets = {names[0]: (0, 1024), names[1]: (1024, 1024), names[2]: (2048, 512), names[3]: (4096, 4000),
                   names[4]: (8128, 64)}
        for step in range(2):
            eng = _fake_engine(total, layer_ranges, rank)
            eng.flat.names, eng.flat.offsets = names, offsets
            expect = sum(_fake_engine(total, layer_ranges, r).flat.G for r in range(world)) / world
            if step == 0:
                sync = GradSync(eng)
                assert sync._emb_lo == 4096 and eng.tail_grad_hook is not None
            else:                                               # same GradSync object, fresh gradients
                sync.engine.flat.G.copy_(eng.flat.G)
                eng = sync.engine
            eng.layer_grad_hook(1)
            eng.layer_grad_hook(0)
            eng.tail_grad_hook()                                 # embedding backward done
            assert sync.tail_ranges() == [(2048, 4096)]          # embeddings no longer part of the tail
            eng.tail_grad_hook()                                 # idempotent within a step
            sync.launch_tail()
            sync.wait_layers()
            sync.wait_tail()
            assert torch.allclose(eng.flat.G, expect, atol=1e-6), "step %d" % step
            assert sync.tail_ranges() == [(2048, total)]         # flag reset: next step reduces them again
        q.put((rank, "ok"))
    except Exception as e:                                       # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gradsync_embedding_hook_world2_gloo():
    from mtvaf_b200 import build
    build.build()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_embeddings, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_gradsync_optimizer_owned_ranges_world2_gloo():
    from mtvaf_b200 import build
    build.build()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_owned, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


@pytest.mark.parametrize("fired", [(2, 1, 0), (2,), ()])
def test_gradsync_world2_gloo(fired):
    from mtvaf_b200 import build
    build.build()                                                # the workers import the package (loads the .so)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fired, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _worker_accum(rank, world, port, q):
    """Gradient accumulation (the reference trainer's gradient_accumulation_steps): micro-batch 1 under no_sync() only
    accumulates, micro-batch 2 reduces the SUM; and two backward passes WITHOUT no_sync (average of averages) give the
    same result, the pre-backward hook waiting for the first pass's all-reduces before the second one accumulates."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mtvaf_b200.optim import GradSync
        layer_ranges = [(0, 512), (512, 1024)]
        total = 2048
        g1 = lambda r: torch.randn(total, generator=torch.Generator().manual_seed(10 + r))
        g2 = lambda r: torch.randn(total, generator=torch.Generator().manual_seed(20 + r))
        expect = sum(g1(r) + g2(r) for r in range(world)) / world
        for use_no_sync in (True, False):
            eng = _fake_engine(total, layer_ranges, rank)
            eng.pre_backward_hook = None
            sync = GradSync(eng)
            assert eng.pre_backward_hook is not None
            eng.flat.G.copy_(g1(rank))                            # "backward" of micro-batch 1
            if use_no_sync:
                with sync.no_sync():
                    eng.pre_backward_hook()
                    for i in (1, 0):
                        eng.layer_grad_hook(i)
                    eng.tail_grad_hook()
                assert not sync.works and not sync.done_layers   # nothing was reduced
            else:
                eng.pre_backward_hook()
                for i in (1, 0):
                    eng.layer_grad_hook(i)
            eng.pre_backward_hook()                              # micro-batch 2 starts: outstanding work is awaited
            eng.flat.G.add_(g2(rank))                            # its wgrads accumulate on top
            for i in (1, 0):
                eng.layer_grad_hook(i)
            sync.finish()
            assert torch.allclose(eng.flat.G, expect, atol=1e-5), use_no_sync
        q.put((rank, "ok"))
    except Exception as e:                                       # pragma: no cover
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_gradsync_gradient_accumulation_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_accum, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_adamw_ranges_follow_reference_groups():
    """modules/train.py:894-926: 'bert' and 'encoder_conv' at args.lr, 'crf'/'fc' at 5e-2, everything else
    (projectors, ANP heads, probe) never updated."""
    from mtvaf_b200 import build
    build.build()
    from mtvaf_b200.optim import reference_groups
    groups = reference_groups(5e-5)

    def hit(name):
        for pred, lr, wd in groups:
            if pred(name):
                return lr, wd
        return None

    assert hit("bert.encoder.layer.3.output.dense.weight") == (5e-5, 1e-2)
    assert hit("encoder_conv.0.weight") == (5e-5, 1e-2)
    assert hit("crf.transitions") == (5e-2, 1e-2)
    assert hit("fc.weight") == (5e-2, 1e-2)
    assert hit("projectors.3.weight") is None
    assert hit("img_classifier.weight") is None
    assert hit("oneWordpsdProbe.oneWordpsdProbe.proj") is None
