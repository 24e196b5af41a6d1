"""TEST INFRASTRUCTURE ONLY -- tests/golden/state_dict_keys.json: the (name, shape) list of `state_dict()` of the
UNMODIFIED reference models, in order.  The reference's checkpoint loaders walk state dicts BY INDEX
(modules/train.py:495-521,928-987) and its optimizer groups select parameters by name substring (:473-483,899-916), so a
drop-in must reproduce names, shapes and order.  Run where /root/reference exists:  python -m oracle.make_state_dict_golden
(the frozen ResNet front-end is bypassed by the feature stub on both sides: `image_model.*` keys are excluded)."""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim                      # noqa: E402
from oracle import mtvaf_oracle as O             # noqa: E402
from oracle.make_golden import hf_config         # noqa: E402


def keys_of(model):
    return [[k, list(v.shape)] for k, v in model.state_dict().items() if not k.startswith("image_model.")]


def main():
    out = {}
    cfg = O.EncoderCfg.roberta_base(vocab_size=1000)          # small vocab: shapes scale, names / order do not
    m2 = ref_shim.build_reference_tvnet2(hf_config(cfg), ref_shim.make_args(), list(range(10)))
    out["TVNetSAModel2/roberta"] = keys_of(m2)
    m1 = ref_shim.build_reference_tvnet2(hf_config(cfg), ref_shim.make_args(vao=False), list(range(10)),
                                         cls_name="TVNetSAModel")
    out["TVNetSAModel/roberta"] = keys_of(m1)
    bcfg = O.EncoderCfg.bert_base(vocab_size=1000)
    mb = ref_shim.build_reference_tvnet2(hf_config(bcfg), ref_shim.make_args(bert_name="bert-base-uncased"),
                                         list(range(10)))
    out["TVNetSAModel2/bert"] = keys_of(mb)
    path = os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")
    json.dump(out, open(path, "w"), indent=0)
    for k, v in out.items():
        print(k, len(v), "entries; first", v[0][0], "last", v[-1][0])


if __name__ == "__main__":
    main()
