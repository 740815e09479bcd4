#!/usr/bin/env python
"""Headline benchmark: Llama-3.1-8B (32 layers, bf16, random-init weights, synthetic prompts) decode
tokens/s and p50 TTFT at TP = --gpus on one B200 box, batch 2, 128-token prompt -> 128 generated tokens
(the shape of the reference's Llama-3.1-8B CI perf test; the reference quotes a 4-layer
truncation at TP=32 on Trn1 — that config is reproduced too and reported under ``ci_4layer``).

    python bench.py --gpus 1 --steps 128 --warmup 8
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 ... bench.py --gpus 8 ...

A "step" is one decode step of the whole batch (2 tokens).  ``value`` is device-timed (CUDA events around K
CUDA-graph replays with on-device token feedback, max over ranks); ``e2e`` goes through the public
``model.forward`` API with host-resident inputs (pinned H2D of ids/positions, D2H of the sampled tokens) every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASELINE_TOK_S = 1665.0   # reference Llama-3.1-8B(4-layer) Trn1 TP=32 e2e throughput
BASELINE_E2E_MS = 306.0

LLAMA31_8B = dict(hidden_size=4096, intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32,
                  num_key_value_heads=8, head_dim=128, vocab_size=128256, max_position_embeddings=131072,
                  rms_norm_eps=1e-5, rope_theta=500000.0, tie_word_embeddings=False,
                  rope_scaling=dict(factor=8.0, high_freq_factor=4.0, low_freq_factor=1.0,
                                    original_max_position_embeddings=8192, rope_type="llama3"))
# test/integration/tp32/models/llama/llama3.1/8b/config.json of the reference
LLAMA31_8B_CI4 = dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=4, num_attention_heads=32,
                      num_key_value_heads=32, head_dim=8, vocab_size=128256, max_position_embeddings=131072,
                      rms_norm_eps=1e-5, rope_theta=500000.0, tie_word_embeddings=True,
                      rope_scaling=dict(factor=16.0, high_freq_factor=2.0, low_freq_factor=1.0,
                                        original_max_position_embeddings=8192, rope_type="llama3"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        hi = [x for x in sm if x > 0.5 * (sm[-1] if sm else 0)]
        med = hi[len(hi) // 2] if hi else (sm[len(sm) // 2] if sm else None)
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """The unmodified reference cannot execute on B200: it imports neuronx_distributed / torch_neuronx /
    neuronxcc (AWS Trainium compiler + runtime) at module import time.  See DESIGN.md "Reference arm"."""
    why = "reference requires AWS Neuron SDK (neuronx_distributed, torch_neuronx, neuronxcc, libnrt) - no CUDA path"
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import neuronx_distributed_inference.models.llama.modeling_llama  # noqa: F401
        why = "reference imported but needs Neuron devices (trace->neuronx-cc->NEFF); cannot run on B200"
    except Exception as e:  # expected
        why = f"reference not runnable on B200: {type(e).__name__}: {str(e)[:120]}"
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"impl": "reference", "unavailable": why}))
    return 0


def build_app(hf_cfg, tp, batch, seq_len, ctx, async_mode, output_logits=False):
    from neuronx_distributed_inference_b200.config import OnDeviceSamplingConfig
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    return build_random_llama(hf_cfg, batch_size=batch, seq_len=seq_len, max_context_length=ctx, device="cuda",
                              tp_degree=tp, dtype="bfloat16", skip_warmup=True, async_mode=async_mode,
                              enable_bucketing=True, output_logits=output_logits,
                              on_device_sampling_config=OnDeviceSamplingConfig(top_k=1))


def device_time_ms(fn, dist_group=None):
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist.is_initialized():
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms


def measure(app, batch, ctx, steps, warmup, n_ttft=9):
    """-> dict(decode ms/step device-timed, ttft p50 ms, e2e tokens/s through the public API, launches/step)"""
    import torch
    from neuronx_distributed_inference_b200 import ops
    dev = app.device
    prompt = torch.randint(0, 100, (batch, ctx))
    mask = torch.ones_like(prompt, dtype=torch.int32)
    # ---- TTFT: prefill of the ctx-token prompt through the public API (H2D of the prompt included)
    ttft = []
    for i in range(n_ttft + 2):
        app.reset()
        ms = device_time_ms(lambda: app(prompt, attention_mask=mask))
        if i >= 2:
            ttft.append(ms)
    ttft.sort()
    first = app(prompt, attention_mask=mask).tokens
    # ---- device-timed decode: CUDA-graph replays with on-device token feedback
    tkg = app.token_generation_model
    feedback_before = tkg.async_feedback
    tkg.async_feedback = True
    g = tkg.graph_for(batch, 1, cur_len=ctx + steps + warmup)
    s0 = dict(ops.stats)
    g.inputs["input_ids"].copy_(first.view(batch, 1))
    g.inputs["position_ids"].fill_(ctx)
    g.inputs["seq_ids"].copy_(torch.arange(batch, dtype=torch.int32))
    torch.cuda.synchronize()
    tkg.replay_steps(g, warmup)
    ms = device_time_ms(lambda: tkg.replay_steps(g, steps))
    tkg.async_feedback = feedback_before
    # my kernels per decode step: count one eager step
    s1 = dict(ops.stats)
    with torch.no_grad():
        app.model(g.inputs["input_ids"], None, g.inputs["position_ids"], g.inputs["seq_ids"], None, is_prefill=False)
    s2 = dict(ops.stats)
    per_step = sum(s2.get(k, 0) - s1.get(k, 0) for k in s2)
    # ---- e2e through the public API: host-resident ids/positions each step, tokens read back each step
    app.reset()
    tok = app(prompt, attention_mask=mask).tokens.cpu()
    pos = torch.full((batch, 1), ctx, dtype=torch.int32)
    pin_ids = torch.empty(batch, 1, dtype=torch.long).pin_memory()
    pin_pos = torch.empty(batch, 1, dtype=torch.int32).pin_memory()
    for _ in range(max(3, warmup)):
        pin_ids.copy_(tok.view(batch, 1)); pin_pos.copy_(pos)
        tok = app(pin_ids, position_ids=pin_pos).tokens.cpu()
        pos += 1

    hist = []

    def e2e_loop():
        nonlocal tok, pos
        for _ in range(steps):
            pin_ids.copy_(tok.view(batch, 1)); pin_pos.copy_(pos)
            tok = app(pin_ids, position_ids=pin_pos).tokens.cpu()   # D2H read of the step result
            hist.append(tok.view(-1).clone())
            pos += 1
    if os.environ.get("NXDI_BENCH_PROFILE_E2E"):
        import cProfile, pstats
        pr = cProfile.Profile(); pr.enable(); e2e_loop(); pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(18)
    e2e_ms = device_time_ms(e2e_loop)
    h2d = pin_ids.numel() * 8 + pin_pos.numel() * 4 + batch * 4 + batch * 12  # ids, positions, seq_ids, sampling params
    d2h = batch * 8
    return dict(ms_per_step=ms / steps, ttft_p50_ms=ttft[len(ttft) // 2], e2e_ms_per_step=e2e_ms / steps,
                launches_per_step=per_step, h2d=h2d, d2h=d2h, checksum=token_checksum(torch.stack(hist[-steps:], 1)))


def token_checksum(tokens):
    """sha1 of the greedy token ids of the timed e2e decode + whether every rank produced the same ids (tensor-parallel
    ranks sample from identical gathered logits: any divergence means a broken collective)."""
    import hashlib
    import torch.distributed as dist
    h = hashlib.sha1(tokens.to("cpu").long().contiguous().numpy().tobytes()).hexdigest()[:16]
    agree = True
    if dist.is_initialized() and dist.get_world_size() > 1:
        allh = [None] * dist.get_world_size()
        dist.all_gather_object(allh, h)
        agree = all(x == allh[0] for x in allh)
    return {"sha1_16": h, "n_tokens": int(tokens.numel()), "ranks_agree": agree,
            "distinct_tokens": int(tokens.unique().numel())}


# ---- the other BASELINE.json configs (bench.py --config ...): same JSON line, measured through the public generate() API ----
DBRX_1L = dict(d_model=6144, n_heads=48, n_layers=1, max_seq_len=32768, vocab_size=100352,
               attn_config=dict(kv_n_heads=8, clip_qkv=8.0, rope_theta=500000.0),
               ffn_config=dict(ffn_hidden_size=10752, moe_num_experts=16, moe_top_k=4, hidden_size=6144))
LLAMA2_7B = dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=32,
                 head_dim=128, vocab_size=32000, max_position_embeddings=4096, rms_norm_eps=1e-5, rope_theta=10000.0)
OPEN_LLAMA_7B = dict(LLAMA2_7B, max_position_embeddings=2048, rms_norm_eps=1e-6)
OPEN_LLAMA_3B = dict(hidden_size=3200, intermediate_size=8640, num_hidden_layers=26, num_attention_heads=32, num_key_value_heads=32,
                     head_dim=100)


def build_config_app(name, tp, batch, seq_len, ctx):
    from neuronx_distributed_inference_b200.config import OnDeviceSamplingConfig
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    common = dict(batch_size=batch, seq_len=seq_len, max_context_length=ctx, device="cuda", tp_degree=tp, dtype="bfloat16",
                  skip_warmup=True, enable_bucketing=True, on_device_sampling_config=OnDeviceSamplingConfig(top_k=1))
    if name == "dbrx":
        from neuronx_distributed_inference_b200.models.dbrx.modeling_dbrx import NeuronDbrxForCausalLM
        return build_random_llama(DBRX_1L, app_cls=NeuronDbrxForCausalLM, **common), \
            "DBRX (1 of 40 layers as in the reference's integration config, 16 experts top-4, random-init)"
    if name == "quant":
        return build_random_llama(LLAMA2_7B, quantized=True, quantization_dtype="f8e4m3", quantization_type="per_channel_symmetric",
                                  **common), "Llama-2-7b (32 layers, per-channel-symmetric fp8-e4m3 weights, random-init)"
    if name == "spec":
        return build_random_llama(OPEN_LLAMA_7B, speculation_length=5, enable_fused_speculation=True,
                                  fused_draft=dict(hf=OPEN_LLAMA_3B, neuron={}), **common), \
            "open_llama_7b target + open_llama_3b draft (fused speculation, k=5, random-init: acceptance is chance level)"
    raise KeyError(name)


def measure_generate(app, batch, ctx, steps, warmup, n_ttft=5):
    """TTFT + tokens/s of ``HuggingFaceGenerationAdapter.generate`` (prompt on the host, sequences read back to the host)."""
    import torch
    from neuronx_distributed_inference_b200 import ops
    from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter
    prompt = torch.randint(1, 100, (batch, ctx))
    mask = torch.ones_like(prompt)
    ttft = []
    for i in range(n_ttft + 2):
        app.reset()
        ms = device_time_ms(lambda: app(prompt, attention_mask=mask.int()))
        if i >= 2:
            ttft.append(ms)
    ttft.sort()
    ad = HuggingFaceGenerationAdapter(app)
    for _ in range(max(1, warmup // 4)):
        app.reset()
        ad.generate(prompt, attention_mask=mask, max_new_tokens=steps)
    s0 = sum(ops.stats.values())
    out = {}

    def run():
        app.reset()
        out["o"] = ad.generate(prompt, attention_mask=mask, max_new_tokens=steps, return_dict_in_generate=True)
    ms = device_time_ms(run)
    launches = sum(ops.stats.values()) - s0
    seq = out["o"].sequences
    # decode replays CUDA graphs (no Python-level dispatch to count): count the kernels of ONE eager decode step instead
    try:
        s1 = sum(ops.stats.values())
        with torch.no_grad():
            app.model(seq[:, -1:].to(app.device), None, torch.full((batch, 1), seq.shape[1] - 1, dtype=torch.int32, device=app.device),
                      torch.arange(batch, dtype=torch.int32, device=app.device), None, is_prefill=False)
        launches = max(launches, (sum(ops.stats.values()) - s1) * int(seq.shape[1] - ctx))
    except Exception:
        pass
    new = seq[:, ctx:]
    stats = getattr(out["o"], "speculation_stats", None)
    return dict(ms_total=ms, new_tokens=int(new.shape[1]), ttft_p50_ms=ttft[len(ttft) // 2], launches=launches,
                checksum=token_checksum(new), speculation_stats=stats, h2d=prompt.numel() * 8 + mask.numel() * 8, d2h=seq.numel() * 8)


def ci_harness(app, batch, ctx, seq_len, n_runs=5, output_logits=False):
    """The reference's benchmark_sampling formula on its CI config: e2e latency of prefill(ctx) + decode to seq_len,
    throughput = n_runs * max_length * batch / total_time (utils/benchmark.py:496-511)."""
    import torch
    prompt = torch.randint(0, 100, (batch, ctx))
    mask = torch.ones_like(prompt, dtype=torch.int32)
    lat = []
    for i in range(n_runs + 1):
        app.reset()
        t0 = time.perf_counter()
        out = app(prompt, attention_mask=mask)
        tok = out.tokens.cpu()
        if output_logits:
            _ = out.logits.cpu()        # the reference's config returns the logits of every step to the host
        pos = torch.full((batch, 1), ctx, dtype=torch.int32)
        for _ in range(seq_len - ctx - 1):
            out = app(tok.view(batch, 1), position_ids=pos)
            tok = out.tokens.cpu()
            if output_logits:
                _ = out.logits.cpu()
            pos += 1
        torch.cuda.synchronize()
        if i > 0:
            lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    p50 = lat[len(lat) // 2]
    thr = len(lat) * seq_len * batch / (sum(lat) / 1e3)
    return dict(model="llama3.1-8b 4-layer CI config (head_dim 8, 32 kv heads, tied embeddings)", e2e_p50_ms=p50,
                throughput_tok_s=thr, vs_baseline_throughput=thr / BASELINE_TOK_S,
                vs_baseline_latency=BASELINE_E2E_MS / p50, timing="host wall-clock like the reference harness",
                output_logits=output_logits)


def _hard_exit(code=0):
    """Tearing down NCCL communicators that were captured into CUDA graphs can block forever in
    destroy_process_group(); results are already printed, so flush and leave."""
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(code)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--ctx", type=int, default=128)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--skip-ci", action="store_true")
    ap.add_argument("--config", default="llama8b", choices=["llama8b", "dbrx", "quant", "spec"],
                    help="BASELINE.json config: llama8b (headline, default), dbrx (1-layer MoE), quant (Llama-2-7b fp8), "
                         "spec (open_llama 7b + 3b draft)")
    ap.add_argument("--shard-shapes", type=int, default=1,
                    help="DIAGNOSTIC (1 GPU): run the per-rank shapes of TP=N without the collectives (compute-only share of a TP step)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    # Libraries (NCCL banner, HF progress bars) print to stdout; the contract is ONE JSON line there.  Route fd 1 to stderr
    # for the whole run and keep the real stdout for the result line.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from neuronx_distributed_inference_b200.parallel import state as pstate
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus or world == 1 and args.gpus == 1, f"launch with torchrun --nproc-per-node {args.gpus}"
    pstate.init_distributed("nccl")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from neuronx_distributed_inference_b200.ops._ext import load_extension
    load_extension()

    seq_len = args.ctx + args.steps + args.warmup + 8
    if args.config != "llama8b":
        return run_other_config(args, rank, real_stdout, seq_len)
    cfg = dict(LLAMA31_8B, num_hidden_layers=args.layers)
    if args.shard_shapes > 1:
        n = args.shard_shapes
        cfg.update(num_attention_heads=32 // n, num_key_value_heads=max(1, 8 // n), intermediate_size=14336 // n,
                   vocab_size=128256 // n)
        args.skip_ci = True
    app = build_app(cfg, args.gpus, args.batch, seq_len, args.ctx, async_mode=False)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    m = measure(app, args.batch, args.ctx, args.steps, args.warmup)
    clocks = sampler.stop()
    tok_s = args.batch * 1e3 / m["ms_per_step"]
    e2e_tok_s = args.batch * 1e3 / m["e2e_ms_per_step"]
    ci = None
    app4 = None
    weights_gb = None
    try:
        weights_gb = sum(p.numel() * p.element_size() for p in app.model.parameters()) / 1e9
    except Exception:
        pass
    if not args.skip_ci and args.layers == 32 and args.gpus > 1:
        # the like-for-like reproduction of the reference's CI configuration is a ONE-GPU block (the reference's number is a fixed
        # TP=32 Trn1 figure, not a scaling curve); a failure of a secondary block on one rank must not be able to wedge the job
        ci = {"skipped": "reported by the 1-GPU run only"}
    elif not args.skip_ci and args.layers == 32:
        # The secondary block must never cost the headline line: any failure in it is reported inside the JSON instead.
        try:
            del app
            torch.cuda.empty_cache()
            tpg = pstate.get_tensor_model_parallel_group()
            tpg.symm = None
            # the published 306 ms / 1665 tok/s were measured with output_logits=True: that is the headline
            # like-for-like block; the logits-off variant is reported next to it
            app4 = build_app(LLAMA31_8B_CI4, args.gpus, 2, 256, 128, async_mode=False, output_logits=True)
            ci = ci_harness(app4, 2, 128, 256, output_logits=True)
            del app4
            torch.cuda.empty_cache()
            tpg.symm = None
            app4 = build_app(LLAMA31_8B_CI4, args.gpus, 2, 256, 128, async_mode=False, output_logits=False)
            ci["logits_off"] = {k: v for k, v in ci_harness(app4, 2, 128, 256, output_logits=False).items()
                                if k in ("e2e_p50_ms", "throughput_tok_s")}
        except Exception as e:      # noqa: BLE001
            ci = dict(ci or {}, error=f"{type(e).__name__}: {str(e)[:200]}")
    out = {
        "metric": "llama3.1-8b_decode_tokens_per_sec" if args.shard_shapes == 1 else
                  f"DIAGNOSTIC_tp{args.shard_shapes}_rank_shapes_no_collectives_tokens_per_sec", "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": tok_s / BASELINE_TOK_S, "dtype": "bf16", "data": "synthetic",
        "impl": "ours", "ttft_p50_ms": m["ttft_p50_ms"],
        "config": {"model": f"Llama-3.1-8B ({args.layers} layers, random-init)", "global_batch": args.batch,
                   "seq_len": seq_len, "prompt_len": args.ctx, "parallelism": f"tp{args.gpus}",
                   "sampling": "on-device greedy (top_k=1)", "kv_cache": "contiguous bf16",
                   "l2_policy": "inputs larger than L2: every step streams all weight shards "
                                "(15 GB / tp >> 126 MB L2); no explicit flush",
                   "baseline_note": "vs_baseline divides by the reference's only published Llama-3.1-8B number "
                                    "(1665 tok/s, 4-LAYER truncation, Trn1 TP=32, harness formula counting prompt "
                                    "tokens); the like-for-like reproduction of that config is under ci_4layer"},
        "e2e": {"value": e2e_tok_s, "unit": "tokens/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                "ms_per_step": m["e2e_ms_per_step"], "path": "NeuronLlamaForCausalLM.forward (pinned H2D ids+positions, "
                "CUDA-graph replay, D2H tokens) per step"},
        "gpu_launches": int(m["launches_per_step"] * args.steps),
        "gpu_launches_per_step": int(m["launches_per_step"]),
        "greedy_checksum": m["checksum"],
        "clocks": clocks,
    }
    if ci is not None:
        out["ci_4layer"] = ci
    if rank == 0:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if dist.is_initialized():
        dist.barrier()
        _hard_exit()
    return 0


def run_other_config(args, rank, real_stdout, seq_len):
    import torch
    import torch.distributed as dist
    seq_len = max(seq_len, args.ctx + 6 * (args.steps + 8))      # room for the speculation windows
    app, desc = build_config_app(args.config, args.gpus, args.batch, seq_len, args.ctx)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    m = measure_generate(app, args.batch, args.ctx, args.steps, args.warmup)
    clocks = sampler.stop()
    decode_ms = max(m["ms_total"] - m["ttft_p50_ms"], 1e-3)
    tok_s = args.batch * m["new_tokens"] * 1e3 / decode_ms
    e2e_tok_s = args.batch * m["new_tokens"] * 1e3 / m["ms_total"]
    out = {"metric": f"{args.config}_generate_decode_tokens_per_sec", "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus,
           "steps": m["new_tokens"], "warmup": args.warmup, "ms_per_step": decode_ms / max(1, m["new_tokens"]), "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "bf16" if args.config != "quant" else "bf16 activations, fp8-e4m3 weights",
           "data": "synthetic", "impl": "ours", "ttft_p50_ms": m["ttft_p50_ms"],
           "config": {"model": desc, "global_batch": args.batch, "seq_len": seq_len, "prompt_len": args.ctx,
                      "parallelism": f"tp{args.gpus}", "sampling": "on-device greedy (top_k=1)",
                      "timing": "device events around HuggingFaceGenerationAdapter.generate (prefill + decode loop, host-resident "
                                "prompt and result); value = new tokens / (total - p50 TTFT)",
                      "l2_policy": "inputs larger than L2: every step streams all weight shards; no explicit flush"},
           "e2e": {"value": e2e_tok_s, "unit": "tokens/s", "h2d_bytes_per_step": m["h2d"] / max(1, m["new_tokens"]),
                   "d2h_bytes_per_step": m["d2h"] / max(1, m["new_tokens"]), "ms_total": m["ms_total"],
                   "path": "HuggingFaceGenerationAdapter.generate"},
           "gpu_launches": int(m["launches"]), "greedy_checksum": m["checksum"], "speculation_stats": m["speculation_stats"],
           "clocks": clocks}
    if rank == 0:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out, default=str) + "\n").encode())
    if dist.is_initialized():
        dist.barrier()
        _hard_exit()
    return 0


if __name__ == "__main__":
    sys.exit(main())
