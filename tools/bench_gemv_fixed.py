"""Fixed cost of one decode GEMV launch inside a CUDA graph (dependent chain y -> x), with / without the RMSNorm prologue."""
import sys
import torch
sys.path.insert(0, ".")
from neuronx_distributed_inference_b200 import ops

dev, dt = "cuda", torch.bfloat16


def run(N, K, T, norm, reps=64, res=False):
    w = (torch.randn(N, K, device=dev) * 0.02).to(dt)
    x = torch.randn(T, K, device=dev, dtype=dt)
    n = torch.ones(K, device=dev, dtype=dt) if norm else None
    square = N == K

    def fn():
        y = x
        for _ in range(reps):
            y2 = ops.linear(y if square else x, w, None, norm_weight=n, norm_eps=1e-5, residual=(y if (res and square) else None))
            y = y2 if square else y
        return y
    for _ in range(2):
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
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 10 / reps
    mb = N * K * 2 / 1e6
    print(f"N={N:6d} K={K:6d} T={T} norm={int(norm)} res={int(res)} dependent={int(square)}: {us:6.2f} us/launch  ({mb:6.1f} MB -> {mb / us / 1e6 * 1e6 / 1e3:5.2f} TB/s)", flush=True)


for T in (2, 8):
    run(16, 4096, T, False)
    run(16, 4096, T, True)
    run(4096, 4096, T, False)
    run(4096, 4096, T, True)
    run(4096, 4096, T, True, res=True)
    run(1536, 4096, T, True)
    run(4096, 1024, T, False)
    run(4096, 3584, T, False)
    run(7168, 4096, T, True)
    run(28672, 4096, T, True)
