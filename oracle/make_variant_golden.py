"""TEST INFRASTRUCTURE ONLY -- tests/golden/tvnet2_variants.pt: the flag branches of TVNetSAModel2.forward
(models/bert_model.py:480-532) run through the UNMODIFIED reference: `noauxloss` (:489), `vao=False` (no ANP heads,
:549-563), `use_probe=False` (bare TokenClassifierOutput, :527-532) and `use_prefix=False` (no visual prompt, :486-492).
Scalars, decoded tags and a gradient fingerprint per variant; inputs / weights come from seeds.
Run where /root/reference exists:  python -m oracle.make_variant_golden"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim                                  # noqa: E402
from oracle import mtvaf_oracle as O                         # noqa: E402
from oracle.make_golden import hf_config, grad_fingerprint   # noqa: E402
from mtvaf_b200 import synthetic as S                        # noqa: E402

VARIANTS = {
    "noauxloss": dict(noauxloss=True),
    "no_vao": dict(vao=False),
    "no_probe": dict(use_probe=False),
    "no_prefix": dict(use_prefix=False, use_probe=True, vao=False),
}
CASE = dict(B=3, L=20, shape="twitter2015", batch_seed=21, param_seed=121, vocab=1200)


def main():
    torch.set_num_threads(8)
    cfg = O.EncoderCfg.roberta_base(vocab_size=CASE["vocab"])
    params = S.init_params(cfg, seed=CASE["param_seed"], ln_jitter=0.05)
    batch = S.make_batch(CASE["B"], CASE["L"], vocab=cfg.vocab_size, shape=CASE["shape"], seed=CASE["batch_seed"])
    gold = {"case": CASE, "variants": {}}
    for name, flags in VARIANTS.items():
        args = ref_shim.make_args(**flags)
        model = ref_shim.build_reference_tvnet2(hf_config(cfg), args, list(range(10)))
        own = model.state_dict()
        model.load_state_dict({k: v for k, v in params.items() if k in own}, strict=False)
        model.eval()
        kw = dict(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                  token_type_ids=batch["token_type_ids"], labels=batch["labels"])
        if args.use_prefix:
            kw.update(images=batch["images"], aux_imgs=batch["aux_imgs"], imagelabel=batch["imagelabel"])
        ret = model(**kw)
        if isinstance(ret, tuple):
            out, prob_loss, img_loss = ret
        else:
            out, prob_loss, img_loss = ret, None, None
        out.loss.backward()
        gold["variants"][name] = {
            "flags": flags, "returns_tuple": isinstance(ret, tuple), "loss": out.loss.detach(),
            "prob_loss": None if prob_loss is None else torch.as_tensor(prob_loss).detach(),
            "img_loss": None if img_loss is None else torch.as_tensor(img_loss).detach(),
            "logits": out.logits,
            "grad_fp": grad_fingerprint([(k, v.grad) for k, v in model.named_parameters()]),
        }
        print(name, "tuple" if isinstance(ret, tuple) else "bare", float(out.loss),
              None if prob_loss is None else float(prob_loss), None if img_loss is None else float(img_loss))
    torch.save(gold, os.path.join(ROOT, "tests", "golden", "tvnet2_variants.pt"))


if __name__ == "__main__":
    main()
