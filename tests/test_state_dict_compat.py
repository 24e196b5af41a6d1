"""CPU: the drop-in modules expose the reference's `state_dict()` -- same names, same shapes, SAME ORDER -- pinned by
tests/golden/state_dict_keys.json, which oracle/make_state_dict_golden.py wrote from the UNMODIFIED reference models.
The reference's checkpoint loaders index state dicts by position (modules/train.py:495-521,928-987) and its optimizer
groups pick parameters by name substring (:473-483,899-916): both silently break if either differs.
No kernels run here (model construction only)."""
import json
import os
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))


def _build(cls_name, kind):
    from mtvaf_b200 import build
    build.build()
    from oracle import mtvaf_oracle as O                    # config helper only
    from oracle.make_golden import hf_config
    from mtvaf_b200 import modules as M
    cfg = O.EncoderCfg.roberta_base(vocab_size=1000) if kind == "roberta" else O.EncoderCfg.bert_base(vocab_size=1000)
    args = SimpleNamespace(bert_name="roberta-base" if kind == "roberta" else "bert-base-uncased", prefix_dim=768,
                           prefix_len=4, use_prefix=True, use_probe=True, beta=0.5, alpha=0.1,
                           vao=(cls_name == "TVNetSAModel2"), noauxloss=False, resnet_root=None, compute_dtype="fp32",
                           num_epochs=30, gcn_layer_number=0, num_layers=0, n_gpu=1)
    return getattr(M, cls_name)(list(range(10)), None, args, config=hf_config(cfg), image_model=M.FeatureStub())


@pytest.mark.parametrize("case", sorted(GOLD))
def test_state_dict_names_shapes_and_order_match_reference(case):
    cls_name, kind = case.split("/")
    m = _build(cls_name, kind)
    got = [[k, list(v.shape)] for k, v in m.state_dict().items() if not k.startswith("image_model.")]
    ref = GOLD[case]
    # buffers that newer transformers versions made non-persistent may be absent on either side
    soft = ("position_ids", "token_type_ids")
    got_f = [e for e in got if not e[0].endswith(soft)]
    ref_f = [e for e in ref if not e[0].endswith(soft)]
    assert [e[0] for e in got_f] == [e[0] for e in ref_f]
    assert got_f == ref_f


def test_optimizer_groups_select_the_same_parameters_as_the_reference():
    """modules/train.py:894-916: 'bert' / 'encoder_conv' names at args.lr, 'crf' / 'fc' at 5e-2; everything else is never
    stepped.  Applied to the REFERENCE's parameter names (golden) the predicates of mtvaf_b200.optim must reproduce that."""
    from mtvaf_b200 import build
    build.build()
    from mtvaf_b200.optim import reference_groups
    groups = reference_groups(5e-5)
    names = [k for k, _ in GOLD["TVNetSAModel2/roberta"]]
    stepped = {n for n in names if any(pred(n) for pred, _, _ in groups)}
    for n in names:
        expect = ("bert" in n) or ("encoder_conv" in n) or ("crf" in n) or n.startswith("fc")
        assert (n in stepped) == expect, n
    assert not any(n.startswith(("projectors.", "img_classifier", "aux_img_classifier", "oneWordpsdProbe")) for n in stepped)
