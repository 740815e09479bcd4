"""GQA sharding plans and the fused QKV / O projections (reference modules/attention/gqa.py).  Implementation: modules/gqa.py."""
from ..gqa import *  # noqa: F401,F403
from ..gqa import GQA, GQAPlan, GroupQueryAttention_O, GroupQueryAttention_QKV, determine_sharding_strategy, get_shardable_head_counts, make_gqa_plan  # noqa: F401
