"""Name-compatible alias of the reference's models/modeling_roberta.py: `from mtvaf_b200.modeling_roberta import RobertaModel`."""
from .modules import RobertaModel  # noqa: F401
