#!/usr/bin/env python
"""Where the bf16 (throughput-mode) error comes from: the same model in fp32 parity mode (pinned to the oracle at 1e-4 by
the GPU tests) and in bf16 mode, layer by layer, plus the head-level figures north_star states tolerances for (logits and
loss <= 2e-2, argmax agreement >= 99.9 %) with the random-init tag head and with a trained-like (ridge-fitted) head.
    python tools/bf16_error_budget.py [B L]   -> prints a table (copy into DESIGN.md section 4)
"""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformers import RobertaConfig                     # noqa: E402
from mtvaf_b200 import synthetic as S                      # noqa: E402
from mtvaf_b200.modules import TVNetSAModel2, FeatureStub  # noqa: E402

DEV = "cuda"


def build(params, dtype, vocab):
    cfg = RobertaConfig(vocab_size=vocab, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                        intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                        pad_token_id=1)
    args = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
                           beta=0.5, alpha=0.1, vao=True, noauxloss=False, resnet_root=None, compute_dtype=dtype,
                           probe_ckpt="")
    m = TVNetSAModel2(list(range(10)), None, args, config=cfg, image_model=FeatureStub())
    m.load_state_dict(params, strict=False)
    return m.to(DEV).eval()


def relmax(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max())


def relrms(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def fit_head(h, labels, mask, lam=20.0):
    """Ridge regression of one-hot tags on the fp32 final hidden states: a stand-in for a TRAINED tag head (emission
    margins O(1) instead of the near-ties of a random-init 768->11 projection)."""
    hm = h[mask].double()
    Y = torch.nn.functional.one_hot(labels[mask], 11).double()
    hc = torch.cat([hm, torch.ones(hm.shape[0], 1, dtype=torch.float64, device=h.device)], 1)
    A = hc.T @ hc + lam * torch.eye(hc.shape[1], dtype=torch.float64, device=h.device)
    Wb = torch.linalg.solve(A, hc.T @ Y)
    return Wb[:-1].T.float().contiguous(), Wb[-1].float().contiguous()


def run(m, batch):
    enc_hs = {}
    out, prob, img = m(**batch)
    return out, prob, img, m.last_emissions.float(), m._last_heads


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    Lq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    vocab = 2000
    from oracle import mtvaf_oracle as O                  # config helper only
    params = S.init_params(O.EncoderCfg.roberta_base(vocab_size=vocab), seed=7, ln_jitter=0.05)
    cpu_batch = S.make_batch(B, Lq, vocab=vocab, shape="twitter2017", seed=8)
    if os.environ.get("MTVAF_LEARNABLE_TASK", "1") == "1":
        # gold tag = function of the token (tests/test_model_gpu.py::_learnable_task_batch): a task a head can learn
        g = torch.Generator().manual_seed(8 + 99)
        lab, mask = cpu_batch["labels"], cpu_batch["attention_mask"]
        width = (vocab - 3) // 11
        cpu_batch["input_ids"] = (3 + width * (lab - 1).clamp_min(0)
                                  + torch.randint(0, width, lab.shape, generator=g)) * mask
    batch = {k: v.to(DEV) for k, v in cpu_batch.items()}
    m32, m16 = build(params, "fp32", vocab), build(params, "bf16", vocab)
    with torch.no_grad():
        kv32, _, _ = m32.get_visual_prompt(batch["images"], batch["aux_imgs"], batch["imagelabel"])
        kv16, _, _ = m16.get_visual_prompt(batch["images"], batch["aux_imgs"], batch["imagelabel"])
        print("prefix K/V: relmax %.2e relrms %.2e" % (relmax(kv16.float(), kv32), relrms(kv16.float(), kv32)))
        e32 = m32.bert(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], past_key_values=kv32,
                       output_hidden_states=True)["hidden_states"]
        e16 = m16.bert(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], past_key_values=kv16,
                       output_hidden_states=True)["hidden_states"]
        for i, (a, b) in enumerate(zip(e16, e32)):
            print("hidden_states[%2d]: relmax %.2e relrms %.2e" % (i, relmax(a.float(), b), relrms(a.float(), b)))
        for head in ("random-init", "ridge-fitted"):
            if head == "ridge-fitted":
                W, bvec = fit_head(e32[12].reshape(-1, 768), batch["labels"].reshape(-1),
                                   batch["attention_mask"].reshape(-1).bool())
                for m in (m32, m16):
                    with torch.no_grad():            # in-place on the Parameter: bumps its version -> bf16 shadow refreshed
                        m.fc.weight.copy_(W)
                        m.fc.bias.copy_(bvec)
            o32, p32, i32, em32, h32 = run(m32, batch)
            o16, p16, i16, em16, h16 = run(m16, batch)
            t32 = [t for s in o32.logits for t in s]
            t16 = [t for s in o16.logits for t in s]
            agree = sum(int(a == b) for a, b in zip(t32, t16)) / len(t32)
            seq_agree = sum(int(a == b) for a, b in zip(list(o32.logits), list(o16.logits))) / B
            top2 = em32.topk(2, -1).values
            margin = (top2[..., 0] - top2[..., 1])[batch["attention_mask"].bool()]
            print("%s head: emissions relmax %.2e relrms %.2e | loss rel %.2e | prob_loss rel %.2e | img rel %.2e | "
                  "tag agreement %.4f (%d tokens), whole-sequence agreement %.4f | fp32 margin min %.3g median %.3g"
                  % (head, relmax(em16, em32), relrms(em16, em32), abs(float(o16.loss) - float(o32.loss)) / abs(float(o32.loss)),
                     abs(float(p16) - float(p32)) / abs(float(p32)), abs(float(i16) - float(i32)) / abs(float(i32)),
                     agree, len(t32), seq_agree, float(margin.min()), float(margin.median())))
            print("   norms relmax %.2e" % relmax(h16["norms"], h32["norms"]))


if __name__ == "__main__":
    main()
