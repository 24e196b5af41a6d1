"""Feature wire format + offline front-end stage (mtvaf_b200/features.py).

CPU: file round trip of the bf16 wire format, header validation, `pyramid_features` through a (random-init) torchvision
ResNet equals what the reference's `ImageModel` produces (oracle/_ref when staged).
GPU: the model fed from the cache (bf16 wire tensors, strided views of one H2D buffer) equals the model fed with the
same features as fp32 tensors rounded to bf16."""
import os
from types import SimpleNamespace

import pytest
import torch

from mtvaf_b200 import synthetic as S
from mtvaf_b200.features import FeatureCache, PYRAMID, pyramid_features


def test_feature_cache_round_trip(tmp_path):
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(7, 4, 3840, 2, 2, generator=g).abs()
    path = os.path.join(tmp_path, "train.mtvf")
    fc = FeatureCache.from_tensor(path, feats)
    assert len(fc) == 7 and fc.n_aux == 3 and os.path.getsize(path) == 4096 + 2 * 7 * 4 * PYRAMID
    rows = fc.batch([5, 0, 5, 2])
    assert rows.shape == (4, 4, 3840, 2, 2) and rows.dtype == torch.bfloat16
    assert torch.equal(rows, feats[[5, 0, 5, 2]].to(torch.bfloat16))
    with open(path, "r+b") as fh:                      # corrupt the header: must be refused, not misread
        fh.write(b"{}")
    with pytest.raises(ValueError):
        FeatureCache.open(path)


def test_pyramid_features_match_reference_image_model(tmp_path):
    """The offline stage computes exactly what the reference's frozen ImageModel feeds get_visual_prompt
    (models/bert_model.py:88-111,536-539), here with a random-init ResNet-50 (no weights in the box)."""
    from mtvaf_b200.modules import ImageModel
    torch.manual_seed(0)
    im = ImageModel(resnet_root=None).eval()
    x = torch.randn(2, 3, 64, 64)
    aux = torch.randn(2, 3, 3, 64, 64)
    f = pyramid_features(im, x, aux)
    assert f.shape == (2, 4, 3840, 2, 2)
    from oracle import ref_shim
    if ref_shim.reference_available():
        R = ref_shim.load_reference()
        ref_im = R.bert_model.ImageModel.__new__(R.bert_model.ImageModel)
        torch.nn.Module.__init__(ref_im)
        ref_im.resnet = im.resnet                      # same frozen weights; the reference's own forward
        main, auxs = ref_im(x, aux)
        want = torch.stack([torch.cat(main, 1)] + [torch.cat(a, 1) for a in auxs], 1)
        assert torch.allclose(f, want, atol=1e-6)
    # and through the cache builder
    fc = FeatureCache.build(os.path.join(tmp_path, "c.mtvf"), im, [(x, aux)], n_samples=2)
    assert torch.equal(fc.batch([0, 1]), f.to(torch.bfloat16))


@pytest.mark.gpu
def test_model_fed_from_wire_format_equals_fp32_features(tmp_path):
    from oracle import mtvaf_oracle as O
    from oracle.make_golden import hf_config
    from mtvaf_b200.modules import TVNetSAModel2, FeatureStub
    dev = torch.device("cuda")
    cfg = O.EncoderCfg.roberta_base(vocab_size=900)
    params = S.init_params(cfg, seed=71, ln_jitter=0.05)
    batch = S.make_batch(4, 32, vocab=900, seed=72)
    feats = torch.cat([batch["images"].unsqueeze(1), batch["aux_imgs"]], 1)           # [B, 4, 3840, 2, 2]
    fc = FeatureCache.from_tensor(os.path.join(tmp_path, "f.mtvf"), feats)
    res = []
    for mode in ("fp32_rounded", "wire"):
        args = SimpleNamespace(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
                               beta=0.5, alpha=0.1, vao=True, noauxloss=False, resnet_root=None, compute_dtype="bf16",
                               probe_ckpt="")
        m = TVNetSAModel2(list(range(10)), None, args, config=hf_config(cfg), image_model=FeatureStub())
        m.load_state_dict(params, strict=False)
        m = m.to(dev).eval()
        b = {k: v.to(dev) for k, v in batch.items()}
        if mode == "wire":
            images, aux = FeatureCache.to_device(fc.batch([0, 1, 2, 3]), dev)
            assert images.dtype == torch.bfloat16 and not images.is_contiguous()
            b["images"], b["aux_imgs"] = images, aux
        else:
            b["images"] = b["images"].to(torch.bfloat16).float()
            b["aux_imgs"] = b["aux_imgs"].to(torch.bfloat16).float()
        out, prob, img = m(**b)
        out.loss.backward()
        res.append((float(out.loss), float(img), m.engine().flat.g("encoder_conv.0.weight").clone()))
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    # (the wgrad kernels accumulate split-K partials with fp32 atomics: equal up to summation order)
    assert float((res[0][2] - res[1][2]).abs().max()) <= 1e-3 * float(res[0][2].abs().max())
