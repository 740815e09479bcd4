"""In-kernel timeline of ONE decode step replayed from its CUDA graph: every decode GEMV / attention CTA records
%globaltimer at entry / exit and clock64 at its phase boundaries (csrc/gemv2.cu, attention.cu `prof`).
Prints, per launch: start / end on the global timeline, and the per-CTA phase durations (median / max):
  pre   = entry -> dependency resolved (griddepcontrol.wait returned): time spent prefetching / waiting
  x     = wait -> activations staged (+norm)          main = tiles streamed + flushed        tail = all-reduce poll / exit
usage: python tools/prof_decode.py [--layers 4] [--shard-shapes 8]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LLAMA31_8B, build_app  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--shard-shapes", type=int, default=1)
    a = ap.parse_args()
    from neuronx_distributed_inference_b200.parallel import state as pstate
    from neuronx_distributed_inference_b200.ops._ext import load_extension
    pstate.init_distributed("nccl")
    torch.cuda.set_device(0)
    C = load_extension()
    cfg = dict(LLAMA31_8B, num_hidden_layers=a.layers)
    if a.shard_shapes > 1:
        n = a.shard_shapes
        cfg.update(num_attention_heads=32 // n, num_key_value_heads=max(1, 8 // n), intermediate_size=14336 // n, vocab_size=128256 // n)
    app = build_app(cfg, 1, 2, 256, 128, False)
    ids = torch.randint(0, 100, (2, 128))
    tok = app(ids, attention_mask=torch.ones_like(ids)).tokens
    tkg = app.token_generation_model
    tkg.async_feedback = True
    buf = torch.zeros(4096 * 148 * 8, dtype=torch.int64, device="cuda")
    C.set_prof(buf)
    g = tkg.graph_for(2, 1, cur_len=200)
    total = C.prof_count()
    C.set_prof(None)
    n = total // 3
    g.inputs["input_ids"].copy_(tok.view(2, 1))
    g.inputs["position_ids"].fill_(128)
    g.inputs["seq_ids"].copy_(torch.arange(2, dtype=torch.int32))
    for _ in range(5):
        g.graph.replay()
    torch.cuda.synchronize()
    clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965
    d = buf.view(-1, 148, 8)[2 * n:3 * n].cpu()
    t_ref = None
    names = ["qkv", "attn", "o", "gate_up", "down"]
    print(f"{n} instrumented launches per step; clock {clk} MHz (clock64 deltas converted with it)")
    print(f"{'#':>3} {'kernel':8} {'ctas':>4} {'start':>8} {'end':>8} | {'pre med':>8} {'pre max':>8} | {'x med':>6} {'x max':>6} | "
          f"{'main med':>8} {'main max':>8} | {'tail max':>8} | units max")
    for i in range(n):
        r = d[i]
        live = r[:, 0] != 0
        if not bool(live.any()):
            continue
        r = r[live]
        gt0, ck0, ck1, ck2, ck3, ck4, gt1, meta = [r[:, k] for k in range(8)]
        if t_ref is None:
            t_ref = int(gt0.min())
        us = lambda x: x.double() / clk
        pre, xs, mainp, tail = us(ck1 - ck0), us(ck2 - ck1), us(ck3 - ck2), us(ck4 - ck3)
        name = names[i % 5] if i < n - 1 else "lm_head"
        units = int((meta >> 32).max())
        print(f"{i:3d} {name:8} {r.shape[0]:4d} {(int(gt0.min()) - t_ref) / 1e3:8.2f} {(int(gt1.max()) - t_ref) / 1e3:8.2f} | "
              f"{pre.median():8.2f} {pre.max():8.2f} | {xs.median():6.2f} {xs.max():6.2f} | {mainp.median():8.2f} {mainp.max():8.2f} | "
              f"{tail.max():8.2f} | {units}")
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
