"""The drop-in modules under the reference's OWN trainer code (north_star: "the MTVAF_training.py / modules/train.py
entry points are kept, so the new path is a drop-in replacement").

The unmodified `modules/train.py` is imported from /root/reference (authoring container) or from the byte-for-byte
copy staged by oracle/make_ref.py under oracle/_ref (GPU box), through oracle/ref_shim.py.

  * GPU: `SATrainer2.multiModal_before_train` (modules/train.py:894-926: torch.optim.AdamW over name-selected groups,
    get_linear_schedule_with_warmup) + `SATrainer2._step` (:859-885) + the loop body of `train` (:618-625) for 3 steps,
    once over the reference's TVNetSAModel2 and once over mtvaf_b200's: same loss trajectory and final weights.
  * CPU: the index-walking checkpoint loaders (`load_pretrained`, `load_pretrained2`, `load_bert`, :928-987 and
    :495-521) applied to a `best_model.pth` written by the drop-in give the same result as for the reference model.
"""
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import ref_shim
from oracle import mtvaf_oracle as O
from oracle.make_golden import hf_config
from mtvaf_b200 import synthetic as S

needs_ref = pytest.mark.skipif(not ref_shim.reference_available(),
                               reason="reference tree not staged (python -m oracle.make_ref)")


def _trainer_args(device, **kw):
    d = dict(device=device, use_prefix=True, use_probe=True, lr=5e-5, warmup_ratio=0.01,
             gradient_accumulation_steps=1, num_epochs=1, local_rank=-1, load_path=None, use_pretrained=False,
             use_152=False, use_101=False, use_34=False, use_18=False, train_batch_size=3, eval_begin_epoch=1)
    d.update(kw)
    return SimpleNamespace(**d)


def _batch_tuple(b):
    """TVSADataset2 item order consumed by SATrainer2._step (modules/train.py:866)."""
    return (b["input_ids"], b["attention_mask"], b["token_type_ids"], b["labels"], torch.zeros_like(b["labels"]),
            b["imagelabel"], b["images"], b["aux_imgs"])


def _drive(T, model, targs, batches, steps_total):
    """The reference's own optimizer setup and step, driven exactly as SATrainer2.train does (:605-625)."""
    tr = T.SATrainer2(model=model, args=targs, label_map={}, logger=None)
    tr.train_num_steps = steps_total
    tr.multiModal_before_train()
    losses = []
    for batch in batches:
        attention_mask, labels, logits, loss, prob_loss, img_loss = tr._step(batch, mode="train")
        loss = loss / targs.gradient_accumulation_steps
        losses.append(float(loss.detach().cpu().item()))
        loss.backward()
        tr.optimizer.step()
        tr.scheduler.step()
        tr.optimizer.zero_grad()
        assert len(logits) == labels.shape[0]                       # List[List[int]] from crf.decode (:511)
        assert len(logits[0]) == int(attention_mask[0].sum())
    return tr, losses


@needs_ref
@pytest.mark.gpu
def test_reference_trainer_drives_dropin_and_matches_reference_model():
    dev = torch.device("cuda")
    T = ref_shim.load_reference_trainer()
    cfg = O.EncoderCfg.roberta_base(vocab_size=1200)
    params = S.init_params(cfg, seed=121, ln_jitter=0.05)
    batches = [_batch_tuple(S.make_batch(3, 20, vocab=cfg.vocab_size, shape="twitter2015", seed=21 + i))
               for i in range(3)]
    # ---- the reference model (PyTorch eager on the same GPU)
    rargs = ref_shim.make_args(device=dev)
    ref = ref_shim.build_reference_tvnet2(hf_config(cfg), rargs, list(range(10)))
    ref.load_state_dict({k: v for k, v in params.items() if k in ref.state_dict()}, strict=False)
    ref.eval()                                   # dropout off on both sides (RNG streams cannot match)
    _, ref_losses = _drive(T, ref, _trainer_args(dev), batches, steps_total=50)
    # ---- the drop-in (fp32 parity mode)
    from mtvaf_b200.modules import TVNetSAModel2, FeatureStub
    margs = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
                            beta=0.5, alpha=0.1, vao=True, noauxloss=False, resnet_root=None, compute_dtype="fp32",
                            device=dev, n_gpu=1)
    m = TVNetSAModel2(list(range(10)), None, margs, config=hf_config(cfg), image_model=FeatureStub())
    m.load_state_dict(params, strict=False)
    m.eval()
    tr, losses = _drive(T, m, _trainer_args(dev), batches, steps_total=50)
    # optimizer groups picked by the reference's name filters over OUR parameter names
    sizes = [len(g["params"]) for g in tr.optimizer.param_groups]
    assert sizes[1] == 4 and sizes[2] == 5 and sizes[0] == len([n for n, _ in m.named_parameters() if "bert" in n])
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 1e-4 * abs(b), (losses, ref_losses)
    rsd = ref.state_dict()
    worst = 0.0
    for k, v in m.state_dict().items():
        if k not in rsd or not v.dtype.is_floating_point:
            continue
        r = rsd[k].float().cpu()
        err = float((v.float().cpu() - r).abs().max() / (r.abs().max() + 1e-12))
        worst = max(worst, err)
        assert err < 2e-4, (k, err)
    print("trainer-driven 3 steps: losses", losses, "reference", ref_losses, "worst weight err %.2e" % worst)


@needs_ref
@pytest.mark.gpu
def test_reference_trainer_drives_dropin_bf16_training_mode():
    """Throughput mode under the reference trainer: torch.optim.AdamW updates the flat fp32 views in place, the bf16
    shadow follows (weights_version), optimizer.zero_grad(set_to_none) re-arms the flat gradient buffer."""
    dev = torch.device("cuda")
    T = ref_shim.load_reference_trainer()
    cfg = O.EncoderCfg.roberta_base(vocab_size=1200)
    params = S.init_params(cfg, seed=122)
    batches = [_batch_tuple(S.make_batch(4, 32, vocab=cfg.vocab_size, seed=40)) for _ in range(6)]
    from mtvaf_b200.modules import TVNetSAModel2, FeatureStub
    margs = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
                            beta=0.5, alpha=0.1, vao=True, noauxloss=False, resnet_root=None, compute_dtype="bf16",
                            device=dev, n_gpu=1)
    m = TVNetSAModel2(list(range(10)), None, margs, config=hf_config(cfg), image_model=FeatureStub())
    m.load_state_dict(params, strict=False)
    m.train()
    w0 = m.fc.weight.detach().clone()
    tr, losses = _drive(T, m, _trainer_args(dev), batches, steps_total=50)
    # (no monotonic-loss assertion: the reference's lr of 5e-2 with Adam on `fc` / `crf` moves every head weight by
    # ~5e-2 per step, i.e. the emissions by O(10): the CRF loss first RISES for tens of steps -- in the reference too;
    # the trajectory itself is pinned against the reference model in the fp32 test above)
    assert all(torch.isfinite(torch.tensor(losses))), losses
    assert not torch.equal(m.fc.weight.detach().cpu(), w0.cpu())
    f = m.engine().flat
    assert f.Wb is not None
    m(**{k: v.to(dev) for k, v in S.make_batch(4, 32, vocab=cfg.vocab_size, seed=40).items()})   # refreshes the shadow
    name = "bert.encoder.layer.0.intermediate.dense.weight"
    assert torch.equal(f.wb(name).float(), f.w(name).to(torch.bfloat16).float())


# ---------------------------------------------------------------------------------------------------------- CPU
def _dropin_cpu(cls_name, cfg, params):
    from mtvaf_b200 import modules as M
    margs = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
                            beta=0.5, alpha=0.1, vao=True, noauxloss=False, resnet_root=None, compute_dtype="fp32",
                            n_gpu=1, num_epochs=30, gcn_layer_number=0, num_layers=0)
    m = getattr(M, cls_name)(list(range(10)), None, margs, config=hf_config(cfg), image_model=M.FeatureStub())
    own = m.state_dict()
    m.load_state_dict({k: v for k, v in params.items() if k in own}, strict=False)
    return m


def _reference_cpu(cls_name, cfg, params, **flags):
    m = ref_shim.build_reference_tvnet2(hf_config(cfg), ref_shim.make_args(**flags), list(range(10)),
                                        cls_name=cls_name)
    own = m.state_dict()
    m.load_state_dict({k: v for k, v in params.items() if k in own}, strict=False)
    return m


def _run_loader(T, trainer_cls, loader, model, path):
    """Outcome of one reference loader: the resulting state_dict, or the exception type it dies with (some
    combinations walk off the end of the key list in the reference itself -- the drop-in must do the same)."""
    tr = getattr(T, trainer_cls)(model=model, args=SimpleNamespace(load_path=path), label_map={}, logger=None)
    try:
        getattr(tr, loader)()
    except Exception as e:                                   # noqa: BLE001 - the outcome IS the exception type
        return type(e).__name__, None
    return "ok", model.state_dict()


@needs_ref
@pytest.mark.parametrize("ckpt_cls,trainer_cls,model_cls,loader", [
    ("TVNetSAModel2", "SATrainer2", "TVNetSAModel2", "load_pretrained2"),
    ("TVNetSAModel2", "SATrainer2", "TVNetSAModel2", "load_bert"),
    ("TVNetSAModel2", "SATrainer2", "TVNetSAModel2", "load_pretrained"),
    ("TVNetSAModel", "SATrainer2", "TVNetSAModel2", "load_pretrained"),
    ("TVNetSAModel", "SATrainer2", "TVNetSAModel2", "load_pretrained2"),
    ("TVNetSAModel2", "SATrainer", "TVNetSAModel", "load_pretrained"),
    ("TVNetSAModel", "SATrainer", "TVNetSAModel", "load_pretrained"),
])
def test_checkpoint_from_dropin_through_reference_index_walking_loaders(tmp_path, ckpt_cls, trainer_cls, model_cls,
                                                                        loader):
    """modules/train.py:495-521 and :928-987 walk two state_dicts BY INDEX (and by 'bert' / 'crf' / 'dense'
    substrings): a `best_model.pth` saved from the drop-in must steer them exactly like one saved from the reference
    model -- same resulting weights, or the same exception where the reference's own walk runs off the list."""
    T = ref_shim.load_reference_trainer()
    cfg = O.EncoderCfg.roberta_base(vocab_size=300)
    trained = S.init_params(cfg, seed=301, ln_jitter=0.05, with_span=True)     # "checkpoint" weights
    fresh = S.init_params(cfg, seed=302, ln_jitter=0.05, with_span=True)       # the model it is loaded into
    vao = dict(vao=False) if ckpt_cls == "TVNetSAModel" else {}
    p_drop, p_ref = os.path.join(tmp_path, "best_model.pth"), os.path.join(tmp_path, "ref_model.pth")
    torch.save(_dropin_cpu(ckpt_cls, cfg, trained).state_dict(), p_drop)
    torch.save(_reference_cpu(ckpt_cls, cfg, trained, **vao).state_dict(), p_ref)
    sd_d, sd_r = torch.load(p_drop), torch.load(p_ref)
    assert list(sd_d.keys()) == list(sd_r.keys())                              # names AND order
    for k in sd_r:
        assert sd_d[k].shape == sd_r[k].shape and torch.equal(sd_d[k].float(), sd_r[k].float()), k
    vao = dict(vao=False) if model_cls == "TVNetSAModel" else {}
    got = _run_loader(T, trainer_cls, loader, _dropin_cpu(model_cls, cfg, fresh), p_drop)
    want = _run_loader(T, trainer_cls, loader, _reference_cpu(model_cls, cfg, fresh, **vao), p_ref)
    assert got[0] == want[0], (got[0], want[0])
    if want[0] != "ok":
        return
    assert list(got[1].keys()) == list(want[1].keys())
    changed = 0
    for k in want[1]:
        assert torch.equal(got[1][k].float(), want[1][k].float()), (loader, k)
        if k in fresh and want[1][k].dtype.is_floating_point and not torch.equal(want[1][k], fresh[k]):
            changed += 1
    assert changed > 10, "the loader was a no-op"
