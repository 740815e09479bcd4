"""Names the reference exposes as compiler custom calls (modules/custom_calls.py:8-45): ``CustomRMSNorm`` (fp32-compute RMSNorm
lowered to ``AwsNeuronRmsNorm``) and ``neuron_cumsum``.  Here they are the engine's RMSNorm kernel and a plain cumsum (the
sampling kernel folds its cumulative sum into the top-k pass)."""
import torch

from .norm import RMSNorm as CustomRMSNorm  # noqa: F401


def neuron_cumsum(x: torch.Tensor, dim: int = -1) -> torch.Tensor:
    return torch.cumsum(x.float(), dim).to(x.dtype)
