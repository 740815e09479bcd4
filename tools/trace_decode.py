"""Kernel timeline of ONE decode step replayed from its CUDA graph (CUPTI via torch.profiler): name, start, duration,
gap to the previous kernel.  usage: [torchrun ...] python tools/trace_decode.py [--gpus N] [--layers L]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LLAMA31_8B, build_app  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--out", default="gpurun_out/trace.json")
    ap.add_argument("--shard-shapes", type=int, default=1, help="1 GPU: per-rank shapes of TP=N, no collectives")
    a = ap.parse_args()
    from neuronx_distributed_inference_b200.parallel import state as pstate
    pstate.init_distributed("nccl")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    rank = int(os.environ.get("RANK", "0"))
    cfg = dict(LLAMA31_8B, num_hidden_layers=a.layers)
    if a.shard_shapes > 1:
        n = a.shard_shapes
        cfg.update(num_attention_heads=32 // n, num_key_value_heads=max(1, 8 // n), intermediate_size=14336 // n, vocab_size=128256 // n)
    app = build_app(cfg, a.gpus, 2, 256, 128, False)
    ids = torch.randint(0, 100, (2, 128))
    tok = app(ids, attention_mask=torch.ones_like(ids)).tokens
    tkg = app.token_generation_model
    tkg.async_feedback = True
    g = tkg.graph_for(2, 1, cur_len=200)
    g.inputs["input_ids"].copy_(tok.view(2, 1))
    g.inputs["position_ids"].fill_(128)
    g.inputs["seq_ids"].copy_(torch.arange(2, dtype=torch.int32))
    for _ in range(5):
        g.graph.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            g.graph.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    rows = [(e.name[:70], e.time_range.start, e.time_range.end - e.time_range.start) for e in evs]
    n = len(rows) // 3
    step = rows[n:2 * n]
    t0 = step[0][1]
    out, prev_end = [], t0
    for name, s, d in step:
        out.append(dict(name=name, start_us=round(s - t0, 2), dur_us=round(d, 2), gap_us=round(s - prev_end, 2)))
        prev_end = max(prev_end, s + d)
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=0)
        total = out[-1]["start_us"] + out[-1]["dur_us"]
        print(f"step span {total:.1f} us, {len(out)} kernels")
        agg = {}
        for r in out:
            k = r["name"].split("(")[0][-48:]
            a0 = agg.setdefault(k, [0, 0.0, 0.0])
            a0[0] += 1; a0[1] += r["dur_us"]; a0[2] += max(r["gap_us"], 0)
        for k, (c, d, gp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{c:4d} x {d / c:8.2f} us  gap-before avg {gp / c:6.2f} us   {k}")
        print("--- first layer timeline")
        for r in out[:40]:
            print(f"{r['start_us']:9.2f} +{r['dur_us']:7.2f}  gap {r['gap_us']:6.2f}  {r['name'][:60]}")
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
