"""KV cache for decoders with vision cross-attention layers (Mllama): the ordinary self-attention cache plus, for every
cross-attention layer, the projected vision K/V of each cache line — written ONCE at prefill, read at every decode step.

reference: modules/kvcache/multimodal_kv_cache_manager.py:11-134 (``[B, H, vision_seq, D]`` entries for the cross layers, a
``vision_key_value`` update path that bypasses the positional scatter) and models/mllama/modeling_mllama.py (cross-attention
cache handling inside the traced graph).

Stored per line next to the K/V: the cross-attention mask row of the LAST prompt token, which is what every generated token
inherits (the Hugging Face / reference behaviour).  Vision length is only known at the first image, so the buffers are allocated
lazily and re-allocated when a request arrives with a different number of vision tokens."""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.nn as nn

from .kv_cache_manager import KVCacheManager


class VisionKVStore(nn.Module):
    def __init__(self, num_lines: int):
        super().__init__()
        self.num_lines = num_lines
        self.k = self.v = self.row_mask = None
        self.has_vision = False

    def reset(self):
        self.has_vision = False

    def store(self, lines: torch.Tensor, k: torch.Tensor, v: torch.Tensor, row_mask: torch.Tensor):
        """k/v [B, Hkv, Nv, D]; row_mask [B, Nv] bool (visibility of the vision tokens for tokens generated later)."""
        B, Hkv, Nv, D = k.shape
        if self.k is None or self.k.shape[2] != Nv or self.k.dtype != k.dtype or self.k.device != k.device:
            self.k = k.new_zeros(self.num_lines, Hkv, Nv, D)
            self.v = v.new_zeros(self.num_lines, Hkv, Nv, D)
            self.row_mask = torch.zeros(self.num_lines, Nv, dtype=torch.bool, device=k.device)
        li = lines.long().clamp(0, self.num_lines - 1)
        self.k[li], self.v[li], self.row_mask[li] = k, v, row_mask
        self.has_vision = True

    def load(self, lines: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        li = lines.long().clamp(0, self.num_lines - 1)
        return self.k[li], self.v[li], self.row_mask[li]

    def bytes(self) -> int:
        return 0 if self.k is None else 2 * self.k.numel() * self.k.element_size() + self.row_mask.numel()


class MultimodalKVCacheManager(KVCacheManager):
    def __init__(self, *a, cross_attention_layers: Iterable[int] = (), **kw):
        super().__init__(*a, **kw)
        lines = self.num_lines + self.garbage
        self.vision = nn.ModuleDict({str(i): VisionKVStore(lines) for i in cross_attention_layers})

    def vision_store(self, layer: int) -> VisionKVStore:
        return self.vision[str(layer)]

    def has_vision(self, layer: int) -> bool:
        return self.vision[str(layer)].has_vision

    def update_vision(self, layer: int, lines, k, v, row_mask):
        self.vision[str(layer)].store(lines, k, v, row_mask)

    def get_vision(self, layer: int, lines):
        return self.vision[str(layer)].load(lines)

    def reset_vision(self):
        for s in self.vision.values():
            s.reset()

    def reset(self):
        super().reset()
        self.reset_vision()

    def bytes(self) -> int:
        return super().bytes() + sum(s.bytes() for s in self.vision.values())
