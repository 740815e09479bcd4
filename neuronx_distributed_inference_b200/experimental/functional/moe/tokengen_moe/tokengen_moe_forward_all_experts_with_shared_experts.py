"""reference experimental/functional/moe/tokengen_moe/tokengen_moe_forward_all_experts_with_shared_experts.py:10."""
from ... import tokengen_moe_megakernel_forward_all_experts_with_shared_experts  # noqa: F401
