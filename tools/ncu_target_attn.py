"""ncu target: tcgen05 flash-attention prefill, B=1 T=2048 Hq=8 Hkv=2 D=128 (3 warm-ups, then the profiled launch)."""
import sys
import torch
sys.path.insert(0, ".")
from neuronx_distributed_inference_b200 import ops
B, T, Hq, Hkv, D = 1, 2048, 8, 2, 128
q = torch.randn(B, T, Hq, D, device="cuda", dtype=torch.bfloat16)
k = torch.randn(B, T, Hkv, D, device="cuda", dtype=torch.bfloat16)
v = torch.randn_like(k)
for _ in range(4):
    o = ops.attention_prefill(q, k, v, D ** -0.5, True)
torch.cuda.synchronize()
print("done")
