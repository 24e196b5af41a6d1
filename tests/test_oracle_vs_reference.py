"""CPU, authoring container only: the oracle against the UNMODIFIED reference imported live from
/root/reference through oracle/ref_shim.py (skipped where the reference is absent, e.g. on the GPU box --
there the committed golden vectors of tests/test_oracle_golden.py pin the oracle instead)."""
import pytest
import torch

from oracle import ref_shim
from oracle import mtvaf_oracle as O
from oracle.make_golden import hf_config
from mtvaf_b200 import synthetic as S

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")]


def test_tvnet2_live_reference_forward_backward():
    cfg = O.EncoderCfg.roberta_base(vocab_size=2000)
    params = S.init_params(cfg, seed=21, ln_jitter=0.05)
    batch = S.make_batch(2, 24, vocab=cfg.vocab_size, shape="twitter2017", seed=22)
    model = ref_shim.build_reference_tvnet2(hf_config(cfg), ref_shim.make_args(), list(range(10)))
    missing = model.load_state_dict(params, strict=False)
    assert not missing.unexpected_keys
    model.eval()
    out, prob_loss, img_loss = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                                     token_type_ids=batch["token_type_ids"], labels=batch["labels"],
                                     imagelabel=batch["imagelabel"], images=batch["images"],
                                     aux_imgs=batch["aux_imgs"])
    out.loss.backward()
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in params.items()}
    o = O.tvnet2_forward(p, cfg, batch, alpha=0.1, beta=0.5)
    o["loss"].backward()
    torch.testing.assert_close(o["loss"], out.loss.detach().cpu(), rtol=1e-5, atol=1e-5)
    # (the reference's probe moves norms / labels to "cuda:0" whenever a GPU exists, probes/probe_trainModel.py:16-17)
    torch.testing.assert_close(o["prob_loss"], prob_loss.detach().cpu(), rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(o["img_loss"], img_loss.detach().cpu(), rtol=1e-5, atol=1e-6)
    assert o["logits"] == out.logits
    ref_grads = dict(model.named_parameters())
    for k in ("fc.weight", "crf.transitions", "encoder_conv.0.weight", "bert.encoder.layer.0.attention.self.query.weight",
              "bert.encoder.layer.11.output.dense.weight", "bert.embeddings.word_embeddings.weight"):
        g_ref, g = ref_grads[k].grad.cpu(), p[k].grad
        assert g_ref is not None and g is not None, k
        n = float(g_ref.norm())
        assert float((g - g_ref).norm()) <= 5e-4 * n + 1e-7, k


def test_construct_label_live_reference():
    R = ref_shim.load_reference()
    g = torch.Generator().manual_seed(9)
    x = torch.rand(5, 64, generator=g) * 30
    x[0, 3] = x[0, 40]
    ref = R.label.ConstructLabelGaget(None)(x)
    assert torch.equal(O.construct_label(x), ref)


@pytest.mark.gpu
def test_oracle_vs_staged_reference_on_the_gpu_box():
    """The same live comparison on the GPU box, against the copy staged by oracle/make_ref.py (oracle/_ref travels
    with the gpurun snapshot; /root/reference does not exist there)."""
    test_tvnet2_live_reference_forward_backward()
    test_construct_label_live_reference()
