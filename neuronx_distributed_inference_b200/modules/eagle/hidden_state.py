"""Rolling buffers that carry target-model state between fused-speculation steps.

reference: modules/eagle/hidden_state.py:75-161 (``HiddenStateRollingBuffer``: a ``[max_bs+1, 2k, H]`` device buffer indexed
by ``(seq_id, position % 2k)`` with an NKI scatter kernel :8-72).  EAGLE drafts consume the *target's* hidden state of the
previous positions; between two fused steps a sequence advances by 1..k positions and the next step looks back at most k,
so a ring of ``2k`` slots per sequence is enough.  Row ``max_bs`` is a garbage row for masked ``seq_ids``.

The scatter is ``index_put_`` on a flattened ``[(max_bs+1)*2k, H]`` view — a single kernel, CUDA-graph capturable."""
from __future__ import annotations

import torch
import torch.nn as nn


class HiddenStateRollingBuffer(nn.Module):
    def __init__(self, max_batch_size: int, buffer_length: int, hidden_size: int, dtype=torch.bfloat16, device=None):
        super().__init__()
        self.max_batch_size, self.buffer_length, self.hidden_size = max_batch_size, buffer_length, hidden_size
        self.register_buffer("hidden_states", torch.zeros(max_batch_size + 1, buffer_length, hidden_size, dtype=dtype,
                                                          device=device), persistent=False)

    def _flat_index(self, seq_ids: torch.Tensor, position_ids: torch.Tensor) -> torch.Tensor:
        """seq_ids [B], position_ids [B,T] -> flat row index [B,T]; invalid seq ids / negative positions -> garbage row."""
        B, T = position_ids.shape
        s = seq_ids.view(B, 1).long().expand(B, T)
        p = position_ids.long()
        bad = (s < 0) | (s >= self.max_batch_size) | (p < 0)
        s = torch.where(bad, torch.full_like(s, self.max_batch_size), s)
        return s * self.buffer_length + p.clamp_min(0) % self.buffer_length

    def set_state(self, seq_ids: torch.Tensor, position_ids: torch.Tensor, hidden_state: torch.Tensor):
        """hidden_state [B,T,H] stored at (seq_id, position % L)."""
        return self.set_state_(seq_ids, position_ids, hidden_state)

    def set_state_(self, seq_ids, position_ids, hidden_state):
        """One scatter kernel, no host sync (CUDA-graph capturable); rows written twice must carry identical values."""
        idx = self._flat_index(seq_ids, position_ids).reshape(-1)
        self.hidden_states.view(-1, self.hidden_size).index_put_((idx,), hidden_state.reshape(-1, self.hidden_size).to(
            self.hidden_states.dtype))
        return self.hidden_states

    def get_state(self, seq_ids: torch.Tensor, position_ids: torch.Tensor) -> torch.Tensor:
        idx = self._flat_index(seq_ids, position_ids)
        return self.hidden_states.view(-1, self.hidden_size)[idx]

    def reset(self):
        self.hidden_states.zero_()


class TokenRollingBuffer(HiddenStateRollingBuffer):
    """Same ring for token ids (the EAGLE draft re-reads the last k accepted tokens together with their target features)."""

    def __init__(self, max_batch_size: int, buffer_length: int, device=None):
        super().__init__(max_batch_size, buffer_length, 1, torch.int64, device)

    def set_tokens(self, seq_ids, position_ids, tokens):
        return self.set_state_(seq_ids, position_ids, tokens.unsqueeze(-1))

    def get_tokens(self, seq_ids, position_ids):
        return self.get_state(seq_ids, position_ids).squeeze(-1)
