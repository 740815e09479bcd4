"""``build`` (reference experimental/core/functions.py:9-80).  The reference traces a prefill and a decode entry point with fixed shapes and
compiles them; on B200 nothing is compiled — "building" fixes the SHAPES a functional model is called with (one entry per model tag),
optionally captures each entry into a CUDA graph, and records ``reserved_example_inputs`` for the bucketing processor."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


class BuiltModel(torch.nn.Module):
    """A functional model behind static-shape entry points.  ``forward(input_tokens, last_pos, attention_mask)`` must match one tag's
    example shapes exactly (pad with :class:`BucketingProcessor` first), like a compiled artefact."""

    def __init__(self, model: torch.nn.Module, cuda_graphs: bool = False):
        super().__init__()
        self.model = model
        self.cuda_graphs = cuda_graphs
        self.reserved_example_inputs: Dict[str, tuple] = {}
        self._graphs: Dict[str, tuple] = {}

    def trace(self, kwargs: dict, tag: str):
        self.reserved_example_inputs[tag] = (kwargs["tokens"], kwargs["last_pos"], kwargs["attention_mask"])
        return self

    def _tag_for(self, tokens, attention_mask) -> str:
        kind = "prefill" if tokens.shape[1] > 1 else "decode"
        for tag, (t, _, m) in self.reserved_example_inputs.items():
            if tag.startswith(kind) and tuple(t.shape) == tuple(tokens.shape) and tuple(m.shape) == tuple(attention_mask.shape):
                return tag
        raise ValueError(f"no built entry for tokens {tuple(tokens.shape)} / mask {tuple(attention_mask.shape)}; built: "
                         f"{ {k: (tuple(v[0].shape), tuple(v[2].shape)) for k, v in self.reserved_example_inputs.items()} }")

    def reset(self):
        if hasattr(self.model, "reset"):
            self.model.reset()

    @torch.no_grad()
    def forward(self, input_tokens=None, last_pos=None, attention_mask=None, tokens=None):
        input_tokens = tokens if input_tokens is None else input_tokens
        tag = self._tag_for(input_tokens, attention_mask)
        dev = next(self.model.parameters(), torch.empty(0)).device
        if not (self.cuda_graphs and dev.type == "cuda"):
            return self.model.forward(input_tokens, last_pos, attention_mask)
        if tag not in self._graphs:      # static buffers + one captured replay per tag
            st = tuple(x.to(dev).clone() for x in (input_tokens, last_pos, attention_mask))
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.model.forward(*st)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.model.forward(*st)
            self._graphs[tag] = (g, st, out)
        g, st, out = self._graphs[tag]
        for dst, src in zip(st, (input_tokens, last_pos, attention_mask)):
            dst.copy_(src)
        g.replay()
        return out.clone()


def build(model: torch.nn.Module, world_size: int = -1, batch_size: int = 1, sequence_length: int = 1024,
          sequence_length_bucketing: bool = True, sequence_length_buckets: Optional[Dict[str, List[int]]] = None,
          cuda_graphs: bool = False) -> BuiltModel:
    """One prefill entry ``[batch, sequence_length]`` and one decode entry ``[batch, 1]`` against a ``sequence_length`` mask
    (the reference's opinionated input contract: ``tokens``, ``last_pos``, ``attention_mask``)."""
    if isinstance(sequence_length, (list, tuple)):
        raise NotImplementedError("use build_flow.build_for_bucketing_on_seq_len for several sequence lengths")
    if sequence_length_buckets:
        from .build_flow import build_for_bucketing_on_seq_len
        return build_for_bucketing_on_seq_len(model, world_size, batch_size, sequence_length_buckets.get("prefill", [sequence_length]),
                                              sequence_length_buckets.get("decode", [sequence_length]), cuda_graphs=cuda_graphs)
    built = BuiltModel(model, cuda_graphs)
    ones = lambda *s: torch.ones(s, dtype=torch.int32)      # noqa: E731
    built.trace(dict(tokens=ones(batch_size, sequence_length), last_pos=torch.zeros(batch_size, dtype=torch.int32),
                     attention_mask=ones(batch_size, sequence_length)), tag="prefill")
    built.trace(dict(tokens=ones(batch_size, 1), last_pos=torch.zeros(batch_size, dtype=torch.int32),
                     attention_mask=ones(batch_size, sequence_length)), tag="decode")
    return built
