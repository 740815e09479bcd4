"""Asynchronous token generation: the host never waits for step N before launching step N+1.

reference: modules/async_execution.py:1-306 (``AsyncTensorWrapper``, ranked-IO sharded next inputs, ``causal_lm_async_execution``
keeps ``prior_outputs`` and returns the previous step's tokens while the current one runs) and the async branches of
``NeuronBaseForCausalLM._get_model_outputs`` (model_base.py:3549-3735).

On B200 the mechanism is the decode CUDA graph itself: with ``async_mode`` the captured step copies its sampled token and
``position + 1`` into its own static inputs, so consecutive replays form a dependent chain on the stream without any host
involvement.  :class:`AsyncDecodeSession` exposes that as an iterator: ``launch()`` enqueues one more step and a D2H copy of
its token into a pinned ring slot guarded by a CUDA event; ``read()`` returns the oldest outstanding result — the host runs
``depth`` steps ahead of what it has read (the reference runs exactly one step ahead)."""
from __future__ import annotations

from collections import deque
from typing import Optional

import torch


class AsyncDecodeSession:
    def __init__(self, app, depth: int = 2):
        if not app.neuron_config.async_mode:
            raise ValueError("NeuronConfig(async_mode=True) is required")
        self.app, self.depth = app, max(1, depth)
        self.runner = app.token_generation_model
        self.graph = None
        self.pending = deque()
        self.on_gpu = app.device.type == "cuda" and self.runner.use_graphs
        self._slots = []

    def start(self, tokens: torch.Tensor, positions: torch.Tensor, seq_ids: Optional[torch.Tensor] = None,
              sampling_params: Optional[torch.Tensor] = None):
        """tokens [B,1] (last sampled), positions [B,1] (their positions)."""
        B = tokens.shape[0]
        self.B = B
        self.seq_ids = seq_ids if seq_ids is not None else torch.arange(B, dtype=torch.int32)
        self.sampling_params = sampling_params
        if self.on_gpu:
            g = self.runner.graph_for(B, 1, cur_len=None)
            si = g.inputs
            si["input_ids"][:B].copy_(tokens, non_blocking=True)
            si["position_ids"][:B].copy_(positions, non_blocking=True)
            si["seq_ids"][:B].copy_(self.seq_ids, non_blocking=True)
            if B < si["seq_ids"].shape[0]:
                si["seq_ids"][B:].fill_(-1)
            if sampling_params is not None:
                si["sampling_params"][:B].copy_(sampling_params, non_blocking=True)
            self.graph = g
            self._slots = [(torch.empty(B, dtype=torch.long).pin_memory(), torch.cuda.Event()) for _ in range(self.depth + 1)]
            self._i = 0
        else:
            self._tok, self._pos = tokens.clone(), positions.clone()
        for _ in range(self.depth):
            self.launch()
        return self

    def launch(self):
        if self.on_gpu:
            self.runner._symm_even()
            self.graph.graph.replay()
            self.runner.n_launch += 1
            buf, ev = self._slots[self._i % len(self._slots)]
            self._i += 1
            buf.copy_(self.graph.out.tokens.view(-1)[: self.B], non_blocking=True)
            ev.record()
            self.pending.append((buf, ev))
        else:   # CPU path: synchronous, same interface
            out = self.app(self._tok, position_ids=self._pos, seq_ids=self.seq_ids, sampling_params=self.sampling_params)
            self._tok = out.tokens.view(-1, 1).clone()
            self._pos = self._pos + 1
            self.pending.append((self._tok.view(-1).clone(), None))

    def read(self) -> torch.Tensor:
        """Tokens of the oldest outstanding step (blocks only on that step's event) and keeps the pipeline full."""
        buf, ev = self.pending.popleft()
        if ev is not None:
            ev.synchronize()
        out = buf.clone()
        self.launch()
        return out

    def stop(self):
        if self.on_gpu:
            torch.cuda.current_stream().synchronize()
        self.pending.clear()


def causal_lm_async_execution(app, tokens, positions, n_steps: int, seq_ids=None, sampling_params=None, depth: int = 2):
    """Generate ``n_steps`` tokens after ``tokens``; -> [B, n_steps] (host tensor).  Equivalent to calling ``app.forward``
    ``n_steps`` times and feeding each result back, without the per-step host round trip."""
    s = AsyncDecodeSession(app, depth).start(tokens, positions, seq_ids, sampling_params)
    outs = [s.read() for _ in range(n_steps)]
    s.stop()
    return torch.stack(outs, 1)


# ---- reference helper names (modules/async_execution.py:10-187) --------------------------------------------------------------
class AsyncTensorWrapper:
    """A device result plus the CUDA event that marks it ready (the reference wraps ranked XLA outputs the same way so that the
    host can defer the synchronisation)."""

    def __init__(self, tensor: torch.Tensor, event: Optional["torch.cuda.Event"] = None):
        self.tensor, self.event = tensor, event

    def sync_async_result_to_cpu(self) -> torch.Tensor:
        if self.event is not None:
            self.event.synchronize()
        return self.tensor.cpu()

    get = sync_async_result_to_cpu


def is_ranked_io(x) -> bool:
    """Ranked IO = per-rank device-resident inputs (no host staging needed).  Here: any CUDA tensor."""
    return torch.is_tensor(x) and x.is_cuda


def will_hit_bucket_boundary(position_ids: torch.Tensor, buckets, look_ahead: int = 1) -> bool:
    """True when one of the next ``look_ahead`` steps needs a larger sequence bucket than the current one — the async
    pipeline then picks the next bucket ahead of time (``second_fit``) instead of re-launching (reference :172-187)."""
    cur = int(position_ids.max()) + 1
    now = next((b for b in sorted(buckets) if b > cur), None)
    later = next((b for b in sorted(buckets) if b > cur + look_ahead), None)
    return now != later


def execute_model(app, *args, **kwargs):
    """Synchronous single step through the application (reference ``execute_model`` :131-146)."""
    return app(*args, **kwargs)


def execute_model_prefix_caching(app, input_dict: dict, pad_type: str = "first_fit"):
    """One step of a prefix-caching request from the dict a serving engine hands over (reference :73-128): derives
    ``num_queries = full_context_lens - computed_context_lens`` when absent and forwards the paged-cache plumbing
    (``slot_mapping``, ``block_table``, context lengths).  -> (AsyncTensorWrapper around the output tokens / logits, is_device)."""
    d = dict(input_dict)
    if "num_queries" not in d and d.get("full_context_lens") is not None:
        d["num_queries"] = d["full_context_lens"] - d["computed_context_lens"]
    out = app(d["input_ids"], d.get("attention_mask"), d.get("position_ids"), d.get("seq_ids"), d.get("sampling_params"),
              adapter_ids=d.get("adapter_ids"), slot_mapping=d.get("slot_mapping"), block_table=d.get("block_table"),
              full_context_lens=d.get("full_context_lens"), computed_context_lens=d.get("computed_context_lens"))
    res = out.tokens if getattr(out, "tokens", None) is not None else out.logits
    ev = None
    if torch.is_tensor(res) and res.is_cuda:
        ev = torch.cuda.Event()
        ev.record()
    return AsyncTensorWrapper(res, ev), bool(torch.is_tensor(res) and res.is_cuda)
