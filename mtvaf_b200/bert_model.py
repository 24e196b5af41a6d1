"""Name-compatible alias of the reference's models/bert_model.py."""
from .modules import TVNetSAModel, TVNetSAModel2, ImageModel, FeatureStub, CRF, TokenClassifierOutput  # noqa: F401
