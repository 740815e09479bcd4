"""Micro-benchmark: persistent GEMV chain vs the same four GEMVs as separate (PDL-chained) launches, inside a CUDA graph."""
import sys
import torch
sys.path.insert(0, ".")
from neuronx_distributed_inference_b200 import ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H, I, NQ = 4096, 14336 // tp, 6144 // tp
HO = 4096 // tp
dev, dt = "cuda", torch.bfloat16
L = 8
W = [dict(wo=(torch.randn(H, HO, device=dev) * 0.02).to(dt), wgu=(torch.randn(2 * I, H, device=dev) * 0.02).to(dt),
          wd=(torch.randn(H, I, device=dev) * 0.02).to(dt), wq=(torch.randn(NQ, H, device=dev) * 0.02).to(dt)) for _ in range(L)]
n = torch.ones(H, device=dev, dtype=dt)
o = torch.randn(T, HO, device=dev, dtype=dt)
h = torch.randn(T, H, device=dev, dtype=dt)


def separate():
    x = h
    for w in W:
        h1 = ops.linear(o, w["wo"], None, residual=x)
        u = ops.linear(h1, w["wgu"], None, norm_weight=n, norm_eps=1e-5, act="silu_mul")
        x = ops.linear(u, w["wd"], None, residual=h1)
        q = ops.linear(x, w["wq"], None, norm_weight=n, norm_eps=1e-5)
    return x


def chained(nph=4):
    x = h
    for w in W:
        ph = [dict(x=o, w=w["wo"], residual=x), dict(x=0, w=w["wgu"], norm=n, eps=1e-5, act="silu_mul"),
              dict(x=1, w=w["wd"], residual=0), dict(x=2, w=w["wq"], norm=n, eps=1e-5)][:nph]
        ys = ops.gemv_chain(ph)
        x = ys[2] if nph >= 3 else h
    return x


def single_phase_chain():
    x = h
    for w in W:
        h1 = ops.gemv_chain([dict(x=o, w=w["wo"], residual=x)])[0]
        u = ops.gemv_chain([dict(x=h1, w=w["wgu"], norm=n, eps=1e-5, act="silu_mul")])[0]
        x = ops.gemv_chain([dict(x=u, w=w["wd"], residual=h1)])[0]
        q = ops.gemv_chain([dict(x=x, w=w["wq"], norm=n, eps=1e-5)])[0]
    return x


def timeit(fn, name, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / reps / L
    byts = sum(v.numel() * 2 for v in W[0].values())
    print(f"{name:28s} {us:8.1f} us/layer   {byts / us / 1e6:6.2f} TB/s   ({byts / 1e6:.0f} MB/layer)", flush=True)


timeit(separate, "separate gemv2 x4")
timeit(single_phase_chain, "chain kernel, 1 phase x4")
timeit(lambda: chained(4), "chain 4 phases")
timeit(lambda: chained(3), "chain 3 phases (no qkv)")
timeit(lambda: chained(2), "chain 2 phases (o+gu)")
