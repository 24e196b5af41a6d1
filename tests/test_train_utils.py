"""CPU: the sync-free trainer bookkeeping (mtvaf_b200/train_utils.py) gives what the reference's per-step loop gives
(modules/train.py:618-661).  GPU: TVNetSAModel's `reuse_extraction` continues forward() from the extraction() pass."""
from types import SimpleNamespace

import pytest
import torch

from mtvaf_b200.train_utils import LossMeter, TagLog

LABEL_MAP = {"O": 1, "B-POS": 2, "I-POS": 3, "B-NEG": 4, "I-NEG": 5, "B-NEU": 6, "I-NEU": 7, "X": 8, "[CLS]": 9,
             "[SEP]": 10}


def _reference_loop(attention_mask, labels, logits, label_map_in):
    """modules/train.py:627-647, restated for the test (test infrastructure)."""
    label_ids = labels.numpy()
    input_mask = attention_mask.numpy()
    label_map = {idx: label for label, idx in label_map_in.items()}
    label_map[0] = "PAD"
    y_true, y_pred = [], []
    for row, mask_line in enumerate(input_mask):
        true_label, true_predict = [], []
        for column, mask in enumerate(mask_line):
            if column == 0:
                continue
            if mask:
                if label_map[label_ids[row][column]] != "X" and label_map[label_ids[row][column]] != "[SEP]":
                    true_label.append(label_map[label_ids[row][column]])
                    true_predict.append(label_map[logits[row][column]])
            else:
                break
        y_true.append(true_label)
        y_pred.append(true_predict)
    return y_true, y_pred


def test_taglog_matches_reference_bookkeeping():
    g = torch.Generator().manual_seed(1)
    log = TagLog(LABEL_MAP)
    want_t, want_p = [], []
    for step in range(3):
        B, Lq = 5, 12
        lens = torch.randint(3, Lq + 1, (B,), generator=g)
        mask = (torch.arange(Lq).unsqueeze(0) < lens.unsqueeze(1)).long()
        labels = torch.randint(1, 9, (B, Lq), generator=g) * mask
        labels[:, 0] = 9
        labels[torch.arange(B), lens - 1] = 10
        decoded = [torch.randint(1, 11, (int(n),), generator=g).tolist() for n in lens]     # List[List[int]] (:511)
        t, p = _reference_loop(mask, labels, decoded, LABEL_MAP)
        want_t += t
        want_p += p
        log.append(mask, labels, decoded)
    got_t, got_p = log.finalize()
    assert got_t == want_t and got_p == want_p
    assert log.finalize() == ([], [])


def test_loss_meter_window_average():
    m = LossMeter(refresh_step=2)
    assert m.add(loss=torch.tensor(2.0), prob_loss=torch.tensor(4.0), img_loss=0) is None
    out = m.add(loss=torch.tensor(4.0), prob_loss=torch.tensor(8.0), img_loss=torch.tensor(1.0))
    assert out == {"loss": 3.0, "prob_loss": 6.0, "img_loss": 0.5}
    assert m.add(loss=torch.tensor(1.0)) is None


@pytest.mark.gpu
def test_taglog_with_device_decoded_tags_and_reuse_extraction():
    from oracle import mtvaf_oracle as O
    from oracle.make_golden import hf_config
    from mtvaf_b200 import synthetic as S, ops
    from mtvaf_b200.modules import TVNetSAModel, TVNetSAModel2, FeatureStub
    dev = torch.device("cuda")
    cfg = O.EncoderCfg.roberta_base(vocab_size=900)
    # ---- TagLog fed from the drop-in's lazily decoded tags: same lists as reading them eagerly
    params = S.init_params(cfg, seed=81, ln_jitter=0.05)
    args = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True, beta=0.5,
                           alpha=0.1, vao=True, noauxloss=False, resnet_root=None, compute_dtype="fp32", probe_ckpt="")
    m2 = TVNetSAModel2(list(range(10)), None, args, config=hf_config(cfg), image_model=FeatureStub())
    m2.load_state_dict(params, strict=False)
    m2 = m2.to(dev).eval()
    batch = S.make_batch(4, 24, vocab=900, seed=82)
    b = {k: v.to(dev) for k, v in batch.items()}
    with torch.no_grad():
        out, _, _ = m2(**b)
    log = TagLog(LABEL_MAP)
    log.append(b["attention_mask"], b["labels"], out.logits)
    got = log.finalize()
    assert got == _reference_loop(batch["attention_mask"], batch["labels"], [list(r) for r in out.logits], LABEL_MAP)

    # ---- reuse_extraction: forward() continues from the extraction() pass (one encoder forward per step)
    sp = S.init_params(cfg, seed=83, ln_jitter=0.05, with_span=True)
    sbatch = S.make_span_batch(3, 32, M=6, vocab=900, seed=84)
    res = {}
    for reuse in (False, True):
        sargs = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
                                beta=0.5, alpha=0.1, vao=False, noauxloss=False, resnet_root=None, compute_dtype="fp32",
                                num_epochs=30, gcn_layer_number=0, num_layers=0, reuse_extraction=reuse)
        m = TVNetSAModel(list(range(10)), None, sargs, config=hf_config(cfg), image_model=FeatureStub())
        own = m.state_dict()
        m.load_state_dict({k: v for k, v in sp.items() if k in own}, strict=False)
        m = m.to(dev).eval()                                  # dropout off so the two flows are comparable
        d = {k: v.to(dev) for k, v in sbatch.items()}
        n0 = ops.launch_count()
        # the trainer's sequence (modules/train.py:341-429): visual prompt -> extraction -> (CPU span candidates) -> forward
        kv = m.get_visual_prompt(d["images"], d["aux_imgs"])
        pm = torch.cat([torch.ones(3, 16, device=dev), d["attention_mask"].float()], 1)
        s_log, e_log, seq, pl = m.extraction(pm, d["input_ids"], kv, d["token_type_ids"])
        keys = ("input_ids", "attention_mask", "token_type_ids", "start_positions", "end_positions", "span_starts",
                "span_ends", "polarity_labels", "label_masks", "images", "aux_imgs")
        out, prob, tot = m(**{k: d[k] for k in keys})
        out.loss.backward()
        torch.cuda.synchronize()
        res[reuse] = (float(out.loss), s_log.clone(), ops.launch_count() - n0,
                      {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None})
    assert abs(res[True][0] - res[False][0]) <= 1e-6 * abs(res[False][0])
    assert torch.equal(res[True][1], res[False][1])
    assert res[True][2] < 0.8 * res[False][2], (res[True][2], res[False][2])      # one encoder forward fewer
    for k, g in res[False][3].items():
        n = float(g.norm())
        if n > 1e-5:                                       # (binary_affine.bias: a sum that cancels to ~1e-7)
            assert float((res[True][3][k] - g).norm()) <= 1e-4 * n, k
