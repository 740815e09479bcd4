"""Tile shape x split-K sweep of the tcgen05 GEMM at prefill sizes where the weights dominate (M = 256 ... 512): one subprocess per
configuration (the overrides are read once per process).  python tools/sweep_skinny_gemm.py"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [("qkv", 6144, 4096, None), ("o", 4096, 4096, None), ("gate_up", 28672, 4096, "silu_mul"), ("down", 4096, 14336, None)]


def child(M):
    import torch
    from neuronx_distributed_inference_b200 import ops
    out = []
    for name, N, K, act in SHAPES:
        x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            ops.linear(x, w, None, act=act)
        ts = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.linear(x, w, None, act=act); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        out.append(f"{name} {ts[3]:.1f}")
    print("RESULT " + " | ".join(out), flush=True)


if len(sys.argv) > 1:
    child(int(sys.argv[1]))
else:
    for M in (256, 512):
        for tm, bn in ((0, 0), (2, 128), (1, 256), (1, 128), (1, 64)):
            for s in (1, 2, 3, 4):
                if (tm, bn) == (0, 0) and s > 1:
                    continue
                e = dict(os.environ)
                if tm:
                    e.update(NXDI_B200_GEMM_TM=str(tm), NXDI_B200_GEMM_BN=str(bn), NXDI_B200_GEMM_SPLITK=str(s))
                r = subprocess.run([sys.executable, __file__, str(M)], env=e, capture_output=True, text=True)
                res = [l for l in r.stdout.split("\n") if l.startswith("RESULT")]
                print(f"M={M} TM={tm} BN={bn} S={s}: {res[0][7:] if res else r.stderr[-200:]}", flush=True)
