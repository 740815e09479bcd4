"""Micro-benchmarks of the hot kernels against the measured roofline (MEASURED_PEAKS.json).
Usage (on the GPU box): python tools/bench_kernels.py [--json gpurun_out/kernels.json]"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuronx_distributed_inference_b200 import ops  # noqa: E402

PEAKS = {"hbm_gbs": 6571.2, "bf16_tflops": 1640.3}
try:
    PEAKS.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass


def timeit(fn, iters=12, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    rows = []
    shapes = [("qkv tp1", 6144, 4096), ("o tp1", 4096, 4096), ("gate_up tp1", 28672, 4096), ("down tp1", 4096, 14336),
              ("lm_head tp1", 128256, 4096), ("qkv tp8", 768, 4096), ("o tp8", 4096, 512), ("gate_up tp8", 3584, 4096),
              ("down tp8", 4096, 1792), ("lm_head tp8", 16032, 4096)]
    for T in (2, 8):
        for name, N, K in shapes:
            x = torch.randn(T, K, device=dev, dtype=torch.bfloat16)
            w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
            g = torch.ones(K, device=dev, dtype=torch.bfloat16)
            act = "silu_mul" if "gate_up" in name else None
            ms = timeit(lambda: ops.linear(x, w, None, norm_weight=g, norm_eps=1e-5, act=act), flush=flush)
            ms_t = timeit(lambda: torch.nn.functional.linear(x, w), flush=flush)
            gb = N * K * 2 / 1e9
            rows.append(dict(kernel="gemv", shape=name, T=T, N=N, K=K, ms=ms, gbs=gb / ms * 1e3,
                             frac_hbm=gb / ms * 1e3 / PEAKS["hbm_gbs"], cublas_ms=ms_t))
            print(f"gemv {name:14s} T={T} N={N:6d} K={K:5d}  {ms*1e3:8.1f} us  {gb/ms*1e3:7.0f} GB/s "
                  f"({gb/ms*1e3/PEAKS['hbm_gbs']*100:5.1f}% of measured HBM)   cuBLAS {ms_t*1e3:8.1f} us")
    # prefill GEMM on tcgen05 vs cuBLAS
    for (M, N, K, name) in [(256, 6144, 4096, "qkv"), (256, 28672, 4096, "gate_up+swiglu"), (256, 4096, 14336, "down"),
                            (2048, 6144, 4096, "qkv"), (2048, 28672, 4096, "gate_up+swiglu"), (2048, 4096, 14336, "down"),
                            (8192, 8192, 8192, "square")]:
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
        act = "silu_mul" if "swiglu" in name else None
        ms = timeit(lambda: ops.linear(x, w, None, act=act), iters=10)
        ms_t = timeit(lambda: torch.nn.functional.linear(x, w), iters=10)
        fl = 2.0 * M * N * K
        rows.append(dict(kernel="gemm_tcgen05", shape=name, M=M, N=N, K=K, ms=ms, tflops=fl / ms / 1e9,
                         frac_bf16=fl / ms / 1e9 / PEAKS["bf16_tflops"], cublas_ms=ms_t, cublas_tflops=fl / ms_t / 1e9))
        print(f"gemm {name:15s} M={M:5d} N={N:6d} K={K:6d} {ms*1e3:9.1f} us {fl/ms/1e9:7.1f} TFLOP/s "
              f"({fl/ms/1e9/PEAKS['bf16_tflops']*100:5.1f}% of measured cuBLAS peak)   cuBLAS {ms_t*1e3:9.1f} us {fl/ms_t/1e9:7.1f}")
    # decode attention
    for (B, Hq, Hkv, S) in [(2, 32, 8, 256), (2, 32, 8, 4096), (2, 32, 8, 32768), (2, 4, 1, 4096), (32, 32, 8, 4096)]:
        D = 128
        kc = torch.randn(B, Hkv, S, D, device=dev, dtype=torch.bfloat16)
        vc = torch.randn_like(kc)
        q = torch.randn(B, 1, Hq, D, device=dev, dtype=torch.bfloat16)
        pos = torch.full((B, 1), S - 1, device=dev, dtype=torch.int32)
        lines = torch.arange(B, device=dev, dtype=torch.int32)
        ms = timeit(lambda: ops.attention_decode(q, kc, vc, lines, pos, 0.088, seq_hint=S), flush=flush)
        gb = 2 * B * Hkv * S * D * 2 / 1e9
        rows.append(dict(kernel="attn_decode", B=B, Hq=Hq, Hkv=Hkv, S=S, ms=ms, gbs=gb / ms * 1e3,
                         frac_hbm=gb / ms * 1e3 / PEAKS["hbm_gbs"]))
        print(f"attn_decode B={B} Hq={Hq} Hkv={Hkv} S={S:6d} {ms*1e3:8.1f} us {gb/ms*1e3:7.0f} GB/s")
    for (B, T, Hq, Hkv) in [(2, 128, 32, 8), (1, 4096, 32, 8), (1, 16384, 32, 8)]:
        D = 128
        q = torch.randn(B, T, Hq, D, device=dev, dtype=torch.bfloat16)
        k = torch.randn(B, T, Hkv, D, device=dev, dtype=torch.bfloat16)
        v = torch.randn_like(k)
        ms = timeit(lambda: ops.attention_prefill(q, k, v, 0.088), iters=5)
        fl = 4 * B * Hq * T * T * D / 2
        rows.append(dict(kernel="attn_prefill", B=B, T=T, Hq=Hq, Hkv=Hkv, ms=ms, tflops=fl / ms / 1e9,
                         frac_bf16=fl / ms / 1e9 / PEAKS["bf16_tflops"]))
        print(f"attn_prefill B={B} T={T} {ms*1e3:9.1f} us {fl/ms/1e9:7.1f} TFLOP/s")
    for V in (128256, 16032):
        x = torch.randn(2, V, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.argmax(x))
        rows.append(dict(kernel="argmax", V=V, ms=ms))
        print(f"argmax V={V} {ms*1e3:.1f} us")
    if a.json:
        os.makedirs(os.path.dirname(a.json), exist_ok=True)
        json.dump(dict(peaks=PEAKS, rows=rows), open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
