"""TEST INFRASTRUCTURE ONLY -- import-time compatibility shim for the *unmodified* reference.

Lets `oracle/make_golden.py` and `tests/test_oracle_vs_reference.py` import MKMaS-GUET/MTVAF's own
modules from /root/reference (read-only, present only in the authoring container) under the installed
torch 2.11 / transformers 5.5, so that the CPU restatement in `oracle/mtvaf_oracle.py` can be validated
against the reference itself and golden vectors can be generated.  Nothing in the product package
(`mtvaf_b200/`) imports this file.  Each shim item follows SURVEY.md section 8(c) items 1-9.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import torch
from torch import nn

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_ref_root() -> str:
    """MTVAF_REF, else the authoring container's /root/reference, else the byte-for-byte copy staged by
    oracle/make_ref.py under oracle/_ref (git-ignored; travels to the GPU box)."""
    cands = [os.environ.get("MTVAF_REF"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "models")):
            return c
    return "/root/reference"


REF_ROOT = _find_ref_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


def _install_transformers_shims():
    import transformers
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    import transformers.file_utils as fu

    # (1) symbols that moved / vanished (models/modeling_roberta.py:43-48)
    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    if not hasattr(mu, "prune_linear_layer"):
        mu.prune_linear_layer = pu.prune_linear_layer
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        def find_pruneable_heads_and_indices(heads, n_heads, head_size, already_pruned_heads):
            raise NotImplementedError("head pruning is not on the hot path")
        mu.find_pruneable_heads_and_indices = find_pruneable_heads_and_indices

    # (2) docstring decorators that reject 4.x-era kwargs (models/modeling_roberta.py:843-848)
    def _noop_decorator(*args, **kwargs):
        def deco(fn):
            return fn
        return deco
    for name in ("add_code_sample_docstrings", "add_start_docstrings",
                 "add_start_docstrings_to_model_forward", "replace_return_docstrings"):
        setattr(fu, name, _noop_decorator)

    # (3,4,5) PreTrainedModel API drift (models/modeling_roberta.py:826,926,944)
    PreTrainedModel = mu.PreTrainedModel
    if not getattr(PreTrainedModel, "_mtvaf_shimmed", False):
        def init_weights(self):
            # 4.x semantics: apply self._init_weights to every sub-module
            self.apply(self._init_weights)
        PreTrainedModel.init_weights = init_weights

        def get_extended_attention_mask(self, attention_mask, input_shape=None, device=None, dtype=None):
            m = attention_mask[:, None, None, :].to(dtype=torch.float32)
            return (1.0 - m) * -10000.0
        PreTrainedModel.get_extended_attention_mask = get_extended_attention_mask

        def get_head_mask(self, head_mask, num_hidden_layers, is_attention_chunked=False):
            return [None] * num_hidden_layers
        PreTrainedModel.get_head_mask = get_head_mask
        PreTrainedModel._mtvaf_shimmed = True


class _StubCRF(nn.Module):
    """Stand-in with pytorch-crf's parameter names/init; the arithmetic used for parity is
    oracle.mtvaf_oracle.crf_* (pytorch-crf is not installed: SURVEY.md 8(c) 'parity unpinned')."""

    def __init__(self, num_tags, batch_first=False):
        super().__init__()
        self.num_tags = num_tags
        self.batch_first = batch_first
        self.start_transitions = nn.Parameter(torch.empty(num_tags))
        self.end_transitions = nn.Parameter(torch.empty(num_tags))
        self.transitions = nn.Parameter(torch.empty(num_tags, num_tags))
        for p in (self.start_transitions, self.end_transitions, self.transitions):
            nn.init.uniform_(p, -0.1, 0.1)

    def forward(self, emissions, tags, mask=None, reduction="sum"):
        from oracle import mtvaf_oracle as O
        llh = O.crf_log_likelihood(emissions, tags, mask, self.start_transitions,
                                   self.end_transitions, self.transitions)
        if reduction == "mean":
            return llh.mean()
        if reduction == "sum":
            return llh.sum()
        return llh

    def decode(self, emissions, mask=None):
        from oracle import mtvaf_oracle as O
        return O.crf_decode(emissions, mask, self.start_transitions, self.end_transitions,
                            self.transitions)


def _install_stub_modules():
    if "torchcrf" not in sys.modules:
        m = types.ModuleType("torchcrf")
        m.CRF = _StubCRF
        sys.modules["torchcrf"] = m
    if "apex" not in sys.modules:
        apex = types.ModuleType("apex")
        apex.amp = types.ModuleType("apex.amp")
        sys.modules["apex"] = apex
        sys.modules["apex.amp"] = apex.amp
    for name in ("tensorboardX", "seqeval", "seqeval.metrics"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sm = sys.modules["seqeval.metrics"]
    if not hasattr(sm, "classification_report"):
        # only called at epoch end by the trainers (modules/train.py:664); the trainer tests drive `_step` directly
        def classification_report(*a, **k):
            raise NotImplementedError("seqeval is not installed (stub)")
        sm.classification_report = classification_report
        sys.modules["seqeval"].metrics = sm


_REF = None


def load_reference():
    """Import the reference packages; returns a namespace of the modules used by the oracle tests."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_transformers_shims()
    _install_stub_modules()
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    for p in (os.path.join(REF_ROOT, "probes"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the product package also has top-level names 'models'/'modules'? No: it lives under mtvaf_b200/.
    import importlib
    ref_models = importlib.import_module("models.bert_model")
    ref_roberta = importlib.import_module("models.modeling_roberta")
    ref_bert = importlib.import_module("models.modeling_bert")
    ref_probe = importlib.import_module("probe")
    ref_label = importlib.import_module("constructLabel")
    ref_probe_model = importlib.import_module("probe_trainModel")
    ref_loss = importlib.import_module("loss")
    _REF = SimpleNamespace(bert_model=ref_models, roberta=ref_roberta, bert=ref_bert, probe=ref_probe,
                           label=ref_label, probe_model=ref_probe_model, loss=ref_loss)
    return _REF


def load_reference_trainer():
    """modules/train.py of the unmodified reference (SATrainer / SATrainer2: `_step`, `multiModal_before_train`, the
    index-walking checkpoint loaders)."""
    load_reference()
    import importlib
    return importlib.import_module("modules.train")


def make_args(**over):
    """args namespace with every attribute the reference model reads."""
    d = dict(bert_name="roberta-base", prefix_dim=768, prefix_len=4, use_prefix=True, use_probe=True,
             beta=0.5, alpha=0.1, n_gpu=1, device=torch.device("cpu"), num_epochs=30,
             use_152=False, use_101=False, use_34=False, use_18=False, resnet_root=None,
             vao=True, noauxloss=False, gcn_layer_number=0, num_layers=0, do_aug=False)
    d.update(over)
    return SimpleNamespace(**d)


class FeatureStub(nn.Module):
    """Replaces the frozen ResNet (out of scope): takes packed pyramid
    features images [B,3840,2,2], aux_imgs [B,n_aux,3840,2,2] and returns them in the list-of-4
    layout `ImageModel.forward` produces (models/bert_model.py:88-111)."""
    WIDTHS = (256, 512, 1024, 2048)

    def _split(self, x):
        return list(torch.split(x, self.WIDTHS, dim=1))

    def forward(self, x, aux_imgs=None):
        main = self._split(x)
        if aux_imgs is None:
            return main, None
        aux = aux_imgs.permute(1, 0, 2, 3, 4)
        return main, [self._split(aux[i]) for i in range(aux.shape[0])]


def build_reference_tvnet2(config, args, label_list, probe_proj=None, seed=0, cls_name="TVNetSAModel2"):
    """Construct the reference TVNetSAModel2 (or the span variant TVNetSAModel) offline (random-init encoder from
    `config`)."""
    R = load_reference()
    bm = R.bert_model
    torch.manual_seed(seed)
    is_roberta = "roberta" in args.bert_name
    enc_cls = R.roberta.RobertaModel if is_roberta else R.bert.BertModel

    orig_from_pretrained = enc_cls.from_pretrained
    orig_image_model = bm.ImageModel
    orig_load = torch.load

    def fake_from_pretrained(name, *a, **k):
        return enc_cls(config)

    class _Img(FeatureStub):
        def __init__(self, *a, **k):
            super().__init__()

    def fake_load(path, *a, **k):
        if isinstance(path, str) and "psdProbe_base_save" in path:
            real = os.path.join(REF_ROOT, "probes", os.path.basename(path))
            k["weights_only"] = False
            return orig_load(real, *a, **k)
        return orig_load(path, *a, **k)

    enc_cls.from_pretrained = staticmethod(fake_from_pretrained)
    bm.ImageModel = _Img
    torch.load = fake_load
    # probe checkpoint was pickled with module path 'probe_trainModel' (on sys.path via probes/)
    try:
        model = getattr(bm, cls_name)(label_list, None, args)
    finally:
        enc_cls.from_pretrained = orig_from_pretrained
        bm.ImageModel = orig_image_model
        torch.load = orig_load
    if probe_proj is not None and args.use_probe:
        with torch.no_grad():
            model.oneWordpsdProbe.oneWordpsdProbe.proj.copy_(probe_proj)
    return model
