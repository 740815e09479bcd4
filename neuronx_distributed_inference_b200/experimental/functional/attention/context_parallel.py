"""reference experimental/functional/attention/context_parallel.py:14-65."""
from .. import gather_kv_context_parallel, split_input_for_context_parallel  # noqa: F401
