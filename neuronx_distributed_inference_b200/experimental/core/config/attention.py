"""reference experimental/core/config/attention.py:5 — which attention implementation a functional model asks for."""
from dataclasses import dataclass


@dataclass
class AttentionConfig:
    try_using_kernel: bool = True      # B200: the CUDA kernels (flash prefill / split-KV decode); False = the fp32 PyTorch definitions
    cp_degree: int = 1
    dp_degree: int = 1
