"""reference experimental/functional/ffn/mlp.py:28-232."""
from .. import gated_mlp, gated_mlp_fused, gated_mlp_kernel_unreduced  # noqa: F401
