"""reference experimental/functional/attention/output_projection.py:12."""
from .. import o_proj_allreduce, o_proj_kernel_unreduced  # noqa: F401
