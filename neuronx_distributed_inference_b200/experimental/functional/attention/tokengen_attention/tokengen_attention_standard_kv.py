"""reference experimental/functional/attention/tokengen_attention/tokengen_attention_standard_kv.py:20."""
from ... import tokengen_attention_megakernel_standard_kv  # noqa: F401
