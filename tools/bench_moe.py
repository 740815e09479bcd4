"""Micro-benchmark: grouped tcgen05 MoE prefill path vs the per-expert cuBLAS loop (ops/reference.py) on DBRX-shaped experts.

python tools/bench_moe.py            (one GPU)
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuronx_distributed_inference_b200 import ops
from neuronx_distributed_inference_b200.ops import reference as ref


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev, dt = "cuda", torch.bfloat16
    torch.manual_seed(0)
    # (name, tokens, top-k, experts, hidden, intermediate shard)
    shapes = [("dbrx tp8 shard, 2048 tok", 2048, 4, 16, 6144, 1344), ("dbrx tp8 shard, 8192 tok", 8192, 4, 16, 6144, 1344),
              ("mixtral tp1, 2048 tok", 2048, 2, 8, 4096, 14336), ("qwen3-moe tp1, 4096 tok", 4096, 8, 128, 2048, 768),
              ("dbrx tp8 shard, 64 tok", 64, 4, 16, 6144, 1344)]
    for name, N, k, E, H, I in shapes:
        x = torch.randn(N, H, device=dev, dtype=dt)
        wgu = (torch.randn(E, 2 * I, H, device=dev) * 0.02).to(dt)
        wd = (torch.randn(E, H, I, device=dev) * 0.02).to(dt)
        idx = torch.rand(N, E, device=dev).topk(k, dim=-1).indices
        w = torch.rand(N, k, device=dev)
        flops = 2.0 * N * k * (2 * I * H + H * I)
        t_g = timeit(lambda: ops.moe_experts(x, wgu, wd, w, idx, "silu_mul", 0))
        t_r = timeit(lambda: ref.moe_experts(x, wgu, wd, w, idx, "silu_mul", 0), iters=5, warm=1)
        print(f"{name}: grouped tcgen05 {t_g:9.1f} us ({flops / t_g / 1e6:7.1f} TFLOP/s)   per-expert cuBLAS loop {t_r:9.1f} us "
              f"({flops / t_r / 1e6:7.1f} TFLOP/s)   x{t_r / t_g:.2f}", flush=True)


if __name__ == "__main__":
    main()
