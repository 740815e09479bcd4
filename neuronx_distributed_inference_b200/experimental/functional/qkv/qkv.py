"""reference experimental/functional/qkv/qkv.py:12."""
from .. import qkv_kernel, qkv_proj  # noqa: F401
