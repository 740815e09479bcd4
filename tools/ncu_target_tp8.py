"""ncu launch list: the four decode GEMVs of one Llama-3.1-8B layer at TP=8 rank shapes (T=2), 4 rounds.
  ncu --set full --clock-control none --import-source on -k regex:gemv2_kernel -s 12 -c 4 -o gpurun_out/ncu_gemv2_tp8 python tools/ncu_target_tp8.py"""
import sys
import torch
sys.path.insert(0, ".")
from neuronx_distributed_inference_b200 import ops

dev, dt = "cuda", torch.bfloat16
T, H, I = 2, 4096, 1792
x = torch.randn(T, H, device=dev, dtype=dt)
n = torch.ones(H, device=dev, dtype=dt)
wq = (torch.randn(768, H, device=dev) * 0.02).to(dt)
wo = (torch.randn(H, 512, device=dev) * 0.02).to(dt)
wgu = (torch.randn(2 * I, H, device=dev) * 0.02).to(dt)
wd = (torch.randn(H, I, device=dev) * 0.02).to(dt)
a = torch.randn(T, 512, device=dev, dtype=dt)
for _ in range(4):
    q = ops.linear(x, wq, None, norm_weight=n, norm_eps=1e-5)
    h1 = ops.linear(a, wo, None, residual=x)
    u = ops.linear(h1, wgu, None, norm_weight=n, norm_eps=1e-5, act="silu_mul")
    y = ops.linear(u, wd, None, residual=h1)
torch.cuda.synchronize()
print("done")
