"""Small launch list for ncu: the decode GEMVs of Llama-3.1-8B (TP1 shapes), the fused rope+attention decode kernel and one
prefill GEMM, each preceded by warm-up launches.  Usage (one GPU):
  ncu --set full --clock-control none --import-source on -k regex:gemv2_kernel -s 6 -c 2 -o gpurun_out/ncu_gemv2 python tools/ncu_target.py"""
import sys
import torch
sys.path.insert(0, ".")
from neuronx_distributed_inference_b200 import ops

dev, dt = "cuda", torch.bfloat16
T, H, I = 2, 4096, 14336
x = torch.randn(T, H, device=dev, dtype=dt)
n = torch.ones(H, device=dev, dtype=dt)
wq = (torch.randn(6144, H, device=dev) * 0.02).to(dt)
wgu = (torch.randn(2 * I, H, device=dev) * 0.02).to(dt)
wd = (torch.randn(H, I, device=dev) * 0.02).to(dt)
for _ in range(4):
    q = ops.linear(x, wq, None, norm_weight=n, norm_eps=1e-5)                      # launches 0..: qkv
    u = ops.linear(x, wgu, None, norm_weight=n, norm_eps=1e-5, act="silu_mul")     # gate_up (GLU)
    y = ops.linear(u, wd, None, residual=x)                                         # down
torch.cuda.synchronize()
# fused rope + append + decode attention, 256-token context
B, S, L, nq, nkv, D = 2, 512, 3, 32, 8, 128
qkv = torch.randn(B, 1, (nq + 2 * nkv) * D, device=dev, dtype=dt)
kc = torch.randn(L, nkv, S, D, device=dev, dtype=dt)
vc = torch.randn(L, nkv, S, D, device=dev, dtype=dt)
lines = torch.tensor([0, 1], device=dev, dtype=torch.int32)
pos = torch.tensor([[255], [200]], device=dev, dtype=torch.int32)
ang = torch.rand(B, 1, D // 2, device=dev)
for _ in range(4):
    o = ops.rope_attention_decode(qkv, ang.cos().contiguous(), ang.sin().contiguous(), kc, vc, lines, pos, pos, nq, nkv, D, D ** -0.5,
                                  seq_hint=S)
# prefill GEMM (M = 256 tokens)
xp = torch.randn(256, H, device=dev, dtype=dt)
for _ in range(3):
    g = ops.linear(xp, wgu, None, act="silu_mul")
    d = ops.linear(g, wd, None, residual=xp)
torch.cuda.synchronize()
print("done")
