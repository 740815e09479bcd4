"""reference experimental/functional/attention/causal_attention_functions.py:19-147."""
from .. import causal_scaled_dot_product_attention, qkv_proj, scaled_dot_product_attention_kernel  # noqa: F401
