"""Continuous batching on top of the application API (the reference leaves scheduling to vLLM; this is the smallest scheduler that
exercises the same hooks): requests arrive over time, each gets a free KV-cache line (``seq_ids``), new requests are prefilled one at a
time while the running ones keep decoding together, finished requests free their line for the next arrival, idle rows are masked
(``seq_id = -1``).

    python examples/continuous_batching_demo.py            # random-weight tiny Llama on CPU / GPU
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch


@dataclass
class Request:
    rid: int
    prompt: List[int]
    max_new_tokens: int
    arrival_step: int = 0
    line: Optional[int] = None
    output: List[int] = field(default_factory=list)

    @property
    def position(self) -> int:                       # position of the NEXT token to feed
        return len(self.prompt) + len(self.output) - 1


class ContinuousBatcher:
    """One decode batch of ``batch_size`` rows; row r of the batch is whatever request currently owns slot r."""

    def __init__(self, app, eos_token_id: Optional[int] = None):
        self.app, self.eos = app, eos_token_id
        nc = app.neuron_config
        self.batch = nc.tkg_batch_size or nc.batch_size
        self.free = list(range(nc.kv_cache_batch_size))
        self.running: Dict[int, Request] = {}        # cache line -> request
        self.done: List[Request] = []

    def _finish_if_needed(self, r: Request):
        if len(r.output) >= r.max_new_tokens or (self.eos is not None and r.output[-1] == self.eos):
            self.done.append(self.running.pop(r.line))
            self.free.append(r.line)

    def admit(self, r: Request) -> bool:
        """Prefill ``r`` into a free cache line (context encoding runs at batch 1, like the reference's CTE with continuous batching)."""
        if not self.free or len(self.running) >= self.batch:
            return False
        r.line = self.free.pop(0)
        ids = torch.tensor([r.prompt])
        out = self.app(ids, attention_mask=torch.ones_like(ids), seq_ids=torch.tensor([r.line], dtype=torch.int32))
        r.output.append(int(out.tokens.reshape(-1)[0]))
        self.running[r.line] = r
        self._finish_if_needed(r)
        return True

    def step(self):
        """One decode step for every running request; unused rows of the batch are masked."""
        if not self.running:
            return
        reqs = list(self.running.values())
        pad = self.batch - len(reqs)
        tok = torch.tensor([[r.output[-1]] for r in reqs] + [[0]] * pad)
        pos = torch.tensor([[r.position] for r in reqs] + [[0]] * pad, dtype=torch.int32)
        seq = torch.tensor([r.line for r in reqs] + [-1] * pad, dtype=torch.int32)
        out = self.app(tok, position_ids=pos, seq_ids=seq)
        new = out.tokens.reshape(-1).tolist()
        for r, t in zip(reqs, new):
            r.output.append(int(t))
            self._finish_if_needed(r)

    def run(self, requests: List[Request], max_steps: int = 10_000) -> List[Request]:
        pending = sorted(requests, key=lambda r: r.arrival_step)
        step = 0
        while (pending or self.running) and step < max_steps:
            while pending and pending[0].arrival_step <= step and self.admit(pending[0]):
                pending.pop(0)
            self.step()
            step += 1
        return sorted(self.done, key=lambda r: r.rid)


def main():
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    tiny = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512, head_dim=64)
    app = build_random_llama(tiny, batch_size=4, seq_len=128, max_context_length=32, device=dev, dtype="float32" if dev == "cpu" else "bfloat16",
                             is_continuous_batching=True, ctx_batch_size=1, kv_cache_batch_size=4, apply_seq_ids_mask=True)
    g = torch.Generator().manual_seed(0)
    reqs = [Request(i, torch.randint(1, 512, (int(torch.randint(4, 20, (1,), generator=g)),), generator=g).tolist(),
                    int(torch.randint(4, 24, (1,), generator=g)), arrival_step=2 * i) for i in range(10)]
    done = ContinuousBatcher(app).run(reqs)
    for r in done:
        print(f"request {r.rid}: prompt {len(r.prompt):2d} tokens, arrived at step {r.arrival_step:2d}, line {r.line}, generated {len(r.output)} tokens")


if __name__ == "__main__":
    main()
