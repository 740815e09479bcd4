"""reference experimental/functional/attention/tokengen_attention/tokengen_attention_block_kv.py:27."""
from ... import tokengen_attention_megakernel_block_kv  # noqa: F401
