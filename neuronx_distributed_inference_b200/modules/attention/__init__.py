"""Attention package — same layout as the reference (modules/attention/{attention_base,gqa,utils,attention_process_groups}.py)."""
from .attention_base import AttentionBase, AttnMeta  # noqa: F401
from .attention_base import AttentionBase as NeuronAttentionBase  # noqa: F401  (reference class name)
