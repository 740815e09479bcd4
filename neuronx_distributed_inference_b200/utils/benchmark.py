"""Benchmark harness (reference utils/benchmark.py:21-516).

Same report shape as the reference — ``{"e2e_model": {...}, "context_encoding_model": {...},
"token_generation_model": {...}}`` each with ``latency_ms_p50/p90/p95/p99/p100/avg`` and
``throughput = n_runs * max_length * max_batch_size / total_time`` (for the token-generation entry ``max_length``
is replaced by the generated-token count, reference :432-446) — but every number is taken twice:
host wall-clock like the reference (``time.perf_counter`` around the call) and, on CUDA, device time from
CUDA events recorded on the launching stream (max over ranks).  ``benchmark_report.json`` is written to the cwd
(reference utils/constants.py:32).
"""
from __future__ import annotations

import json
import time
from typing import Callable, Dict, List, Optional

import torch

from .hf_adapter import HuggingFaceGenerationAdapter

BENCHMARK_REPORT_FILENAME = "benchmark_report.json"


def _percentile(xs: List[float], q: float) -> float:
    if not xs:
        return float("nan")
    s = sorted(xs)
    k = (len(s) - 1) * q / 100.0
    lo, hi = int(k), min(int(k) + 1, len(s) - 1)
    return s[lo] + (s[hi] - s[lo]) * (k - lo)


def generate_report(latency_list_s: List[float], max_length: int, max_batch_size: int, n_runs: Optional[int] = None,
                    device_ms: Optional[List[float]] = None) -> dict:
    ms = [x * 1e3 for x in latency_list_s]
    total = sum(latency_list_s)
    n_runs = n_runs or len(latency_list_s)
    rep = {f"latency_ms_p{p}": _percentile(ms, p) for p in (50, 90, 95, 99, 100)}
    rep["latency_ms_avg"] = sum(ms) / max(len(ms), 1)
    rep["throughput"] = (n_runs * max_length * max_batch_size) / total if total > 0 else float("nan")
    if device_ms:
        rep["device_latency_ms_p50"] = _percentile(device_ms, 50)
        rep["device_latency_ms_p99"] = _percentile(device_ms, 99)
    return rep


class LatencyCollector:
    """Pre/post hook pair around a sub-model runner (reference :484-493); wall-clock + CUDA events."""

    def __init__(self, use_cuda_events: bool = False):
        self.latency_list: List[float] = []
        self.device_ms: List[float] = []
        self._t0 = None
        self._events = []
        self.use_cuda_events = use_cuda_events

    def pre_hook(self, *args):
        if self.use_cuda_events:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            self._e0 = e0
        self._t0 = time.perf_counter()

    def hook(self, *args):
        if self.use_cuda_events:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self._events.append((self._e0, e1))
        self.latency_list.append(time.perf_counter() - self._t0)

    def finalize(self):
        if self._events:
            torch.cuda.synchronize()
            self.device_ms = [a.elapsed_time(b) for a, b in self._events]
            self._events = []


class Benchmark:
    def __init__(self, benchmark_func: Callable, input_param=None, num_runs: int = 20, preprocess_func=None,
                 post_warmup_func=None):
        self.benchmark_func = benchmark_func
        self.input_param = input_param
        self.num_runs = num_runs
        self.preprocess_func = preprocess_func
        self.post_warmup_func = post_warmup_func
        self.latency_list: List[float] = []

    def run(self):
        self._call()  # warm-up
        if self.post_warmup_func:
            self.post_warmup_func()
        for _ in range(self.num_runs):
            if self.preprocess_func:
                self.preprocess_func()
            t0 = time.perf_counter()
            self._call()
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self.latency_list.append(time.perf_counter() - t0)
        return self.latency_list

    def _call(self):
        p = self.input_param
        if p is None:
            return self.benchmark_func()
        if isinstance(p, (tuple, list)):
            return self.benchmark_func(*p)
        if isinstance(p, dict):
            return self.benchmark_func(**p)
        return self.benchmark_func(p)


def get_sample_inputs(neuron_config, vocab_cap: int = 100):
    """``torch.randint(0,100,(batch, input_length))`` with input_length = max_context_length (or half of it
    when it equals max_length) — reference :210-230."""
    L = neuron_config.max_context_length
    if L == neuron_config.max_length:
        L = L // 2
    ids = torch.randint(0, vocab_cap, (neuron_config.batch_size, L))
    return ids, torch.ones_like(ids)


def create_submodule_latency_collectors(model, use_cuda_events: bool = False) -> Dict[str, "LatencyCollector"]:
    """One collector per sub-model runner (context encoding, token generation, speculation / fused speculation, vision encoder) —
    reference utils/benchmark.py:397-407."""
    return {r.tag: LatencyCollector(use_cuda_events) for r in model.models}


def register_latency_collectors(latency_collectors: Dict[str, "LatencyCollector"], model) -> None:
    """Attach the collectors: every runner call is bracketed by ``pre_hook`` / ``hook`` (reference :409-430 uses module forward hooks;
    the runners call their collector directly so that CUDA-graph replays are timed as well)."""
    for r in model.models:
        r.collector = latency_collectors.get(r.tag)


def unregister_latency_collectors(model) -> None:
    for r in model.models:
        r.collector = None


def generate_submodule_reports(latency_collectors: Dict[str, "LatencyCollector"], model, prompt_len: int) -> Dict[str, dict]:
    """Percentile report per sub-model (reference :432-446).  Tokens per call: the prompt length for context encoding, one step's
    worth for the decode-side runners (their latency list has one entry per step)."""
    nc = model.neuron_config
    reports = {}
    for r in model.models:
        col = latency_collectors.get(r.tag)
        if col is None or not col.latency_list:
            continue
        col.finalize()
        n_tok = prompt_len if r.is_prefill else 1
        reports[r.tag] = generate_report(col.latency_list, n_tok, nc.max_batch_size, len(col.latency_list), col.device_ms)
    return reports


def benchmark_sampling(model, draft_model=None, generation_config=None, target: str = "all", num_runs: int = 20,
                       benchmark_report_path: Optional[str] = BENCHMARK_REPORT_FILENAME, image=None) -> Dict[str, dict]:
    """End-to-end ``generate`` benchmark + per-sub-model latencies.  EOS is disabled (generation always runs to
    ``max_length``) like the reference does under on-device sampling (:56-61,77-87)."""
    nc = model.neuron_config
    ids, mask = get_sample_inputs(nc)
    adapter = HuggingFaceGenerationAdapter(model)
    use_ev = model.device is not None and model.device.type == "cuda"
    report: Dict[str, dict] = {}

    def run_generate():
        return adapter.generate(ids, attention_mask=mask, max_length=nc.max_length, eos_token_id=None)

    collectors = create_submodule_latency_collectors(model, use_ev) if target in ("all", "context_encode", "token_gen", "speculation") else {}
    bench = Benchmark(run_generate, num_runs=num_runs)
    bench.post_warmup_func = lambda: register_latency_collectors(collectors, model)     # hooks go in after warm-up (reference :120-123)
    try:
        lat = bench.run()
    finally:
        unregister_latency_collectors(model)
    report["e2e_model"] = generate_report(lat, nc.max_length, nc.max_batch_size, num_runs)
    report.update(generate_submodule_reports(collectors, model, ids.shape[1]))
    if benchmark_report_path:
        with open(benchmark_report_path, "w") as f:
            json.dump(report, f, indent=2)
    return report
