"""Per-cache-line recurrent state for hybrid decoders (short convolutions, gated linear recurrences, SSMs): fixed-size tensors that
play the role the KV cache plays for attention layers — written at the end of prefill, read-modify-written at every decode step,
addressed by the same cache lines (``seq_ids``) so continuous batching works unchanged.

reference: the contrib hybrid ports keep such states as extra aliased graph inputs/outputs next to the KV cache
(contrib/models/{lfm2-2.6b, recurrentgemma-2b-it, Falcon-H1-0.5B-Instruct}/src).  Here they are plain device buffers with stable
addresses (CUDA-graph friendly); masked rows go to the same garbage line as in ``KVCacheManager``."""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn as nn


class RecurrentStateCache(nn.Module):
    def __init__(self, specs: Dict[str, Sequence[int]], num_lines: int, dtype=torch.float32, device=None):
        """``specs``: state name -> per-line shape, or ``(shape, dtype)`` to override the cache dtype (fp32 SSM / LRU states next to
        bf16 activations).  One garbage line is appended."""
        super().__init__()
        self.num_lines = num_lines
        self._names = {}
        for i, (name, shape) in enumerate(specs.items()):
            buf, dt = f"state_{i}", dtype
            if len(shape) == 2 and isinstance(shape[1], torch.dtype):
                shape, dt = shape
            self._names[name] = buf
            self.register_buffer(buf, torch.zeros(num_lines + 1, *shape, dtype=dt, device=device), persistent=False)

    def _lines(self, lines: torch.Tensor) -> torch.Tensor:
        lines = lines.long()
        return torch.where((lines < 0) | (lines > self.num_lines), torch.full_like(lines, self.num_lines), lines)

    def read(self, name: str, lines: torch.Tensor) -> torch.Tensor:
        return getattr(self, self._names[name])[self._lines(lines)]

    def write(self, name: str, lines: torch.Tensor, value: torch.Tensor):
        buf = getattr(self, self._names[name])
        buf[self._lines(lines)] = value.to(buf.dtype)

    def reset(self):
        for b in self.buffers():
            b.zero_()

    def bytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self.buffers())
