"""reference experimental/functional/norm/norm_functions.py:8."""
from .. import rms_norm


def rmsnorm(x, weight, eps: float = 1e-6):
    return rms_norm(x, weight, eps)
