"""Name-compatible alias of the reference's models/modeling_bert.py: `from mtvaf_b200.modeling_bert import BertModel`."""
from .modules import BertModel  # noqa: F401
