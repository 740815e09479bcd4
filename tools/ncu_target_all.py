"""ncu target: ONE launch of every hot kernel at a representative Llama-3.1-8B / DBRX shape inside a cudaProfiler range.

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/ncu_r2_all python tools/ncu_target_all.py

Warm-ups run outside the range (not profiled).  One GPU only.
"""
import math
import sys
import torch
sys.path.insert(0, ".")
from neuronx_distributed_inference_b200 import ops

dev, dt = "cuda", torch.bfloat16
torch.manual_seed(0)
H, I, V = 4096, 14336, 128256
C = ops._C()

# ---- decode (T = 2 tokens, TP1 shapes)
x2 = torch.randn(2, H, device=dev, dtype=dt)
gam = torch.ones(H, device=dev, dtype=dt)
w_qkv = (torch.randn(6144, H, device=dev) * 0.02).to(dt)
w_gu = (torch.randn(2 * I, H, device=dev) * 0.02).to(dt)
w_dn = (torch.randn(H, I, device=dev) * 0.02).to(dt)
w_lm = (torch.randn(V, H, device=dev) * 0.02).to(dt)
B, S, nq, nkv, D = 2, 512, 32, 8, 128
kc = torch.randn(B + 1, nkv, S, D, device=dev, dtype=dt)
vc = torch.randn(B + 1, nkv, S, D, device=dev, dtype=dt)
lines = torch.arange(B, device=dev, dtype=torch.int32)
pos = torch.full((B, 1), 300, device=dev, dtype=torch.int32)
ang = torch.rand(B, 1, D // 2, device=dev) * 6.28
cos, sin = ang.cos().contiguous(), ang.sin().contiguous()

# ---- prefill (2048 tokens)
M = 2048
xm = torch.randn(M, H, device=dev, dtype=dt)
hm = torch.randn(M, I, device=dev, dtype=dt)
q = torch.randn(1, 4096, 32, D, device=dev, dtype=dt)
k = torch.randn(1, 4096, 8, D, device=dev, dtype=dt)
v = torch.randn_like(k)
wq8 = (torch.randn(H, I, device=dev) / math.sqrt(I)).to(torch.float8_e4m3fn)
ws8 = torch.ones(H, device=dev)

# ---- MoE (DBRX TP8 shard, 2048 tokens)
E, kk, Hm, Im = 16, 4, 6144, 1344
xe = torch.randn(M, Hm, device=dev, dtype=dt)
wgu_e = (torch.randn(E, 2 * Im, Hm, device=dev) * 0.02).to(dt)
wd_e = (torch.randn(E, Hm, Im, device=dev) * 0.02).to(dt)
idx = torch.rand(M, E, device=dev).topk(kk, dim=-1).indices
wts = torch.rand(M, kk, device=dev)
logits = torch.randn(2, V, device=dev, dtype=dt)


def run_all():
    qkv = ops.linear(x2, w_qkv, None, norm_weight=gam, norm_eps=1e-5)                       # gemv2 (fused RMSNorm)
    ops.rope_attention_decode(qkv.view(B, 1, -1), cos, sin, kc, vc, lines, pos, pos, nq, nkv, D, D ** -0.5, seq_hint=S)
    u = ops.linear(x2, w_gu, None, norm_weight=gam, norm_eps=1e-5, act="silu_mul")          # gemv2 GLU
    ops.linear(u, w_dn, None, residual=x2)                                                  # gemv2 + residual
    lg = ops.linear(x2, w_lm, None, norm_weight=gam, norm_eps=1e-5)                         # lm_head
    ops.argmax(lg)
    ops.sample(logits, torch.tensor([50, 50], device=dev), torch.tensor([0.9, 0.9], device=dev), torch.tensor([1.0, 1.0], device=dev))
    ops.linear(xm, w_gu, None, norm_weight=gam, norm_eps=1e-5, act="silu_mul")              # tcgen05 GEMM, GLU epilogue
    ops.linear(hm, w_dn, None, residual=xm)                                                 # tcgen05 GEMM + residual
    ops.attention_prefill(q, k, v, D ** -0.5, True)                                         # tcgen05 flash attention
    hq, hs = C.rmsnorm_quant(hm, None, 1e-5, 0.0, float("inf"))                             # per-token fp8 quant
    C.gemm_fp8(hq, hs, wq8, ws8, None, 0, None)                                             # tcgen05 kind::f8f6f4 GEMM
    ops.moe_experts(xe, wgu_e, wd_e, wts, idx, "silu_mul", 0)                               # plan, gather, 2 grouped GEMMs, combine


for _ in range(3):
    run_all()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run_all()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
