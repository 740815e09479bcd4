"""Tiny driver for ncu --set full captures of the two flagship kernels (one launch each after warm-up)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuronx_distributed_inference_b200 import ops  # noqa: E402

dev = "cuda"
x = torch.randn(2, 4096, device=dev, dtype=torch.bfloat16)
g = torch.ones(4096, device=dev, dtype=torch.bfloat16)
w = (torch.randn(28672, 4096, device=dev) / 64).to(torch.bfloat16)
xm = torch.randn(2048, 4096, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.linear(x, w, None, norm_weight=g, norm_eps=1e-5, act="silu_mul")     # gemv2_kernel<true,0>
    ops.linear(xm, w, None, act="silu_mul")                                  # gemm_tcgen05_kernel
torch.cuda.synchronize()
print("done")
