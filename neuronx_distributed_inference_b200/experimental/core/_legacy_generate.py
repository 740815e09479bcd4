"""Greedy generation for the functional models (reference experimental/core/generate/generate.py:1-110): the model signature is
``forward(input_tokens, last_pos, attention_mask) -> logits or tokens [B]``."""
from __future__ import annotations

import torch


@torch.no_grad()
def generate(model, prompt_tokens: torch.Tensor, attention_mask: torch.Tensor = None, max_new_tokens: int = 16, eos_token_id=None,
             pad_token_id: int = 0):
    B, T = prompt_tokens.shape
    if attention_mask is None:
        attention_mask = torch.ones_like(prompt_tokens)
    last = attention_mask.long().sum(-1) - 1
    model.reset()
    nxt = model.forward(prompt_tokens, last, attention_mask)
    seqs = [prompt_tokens[b, : int(last[b]) + 1].tolist() for b in range(B)]
    done = [False] * B
    for _ in range(max_new_tokens):
        nxt = nxt.view(B).cpu()
        for b in range(B):
            if not done[b]:
                seqs[b].append(int(nxt[b]))
                done[b] = eos_token_id is not None and int(nxt[b]) == eos_token_id
        if all(done):
            break
        last = last + 1
        nxt = model.forward(nxt.view(B, 1), last, None)
    width = max(len(s) for s in seqs)
    out = torch.full((B, width), pad_token_id, dtype=torch.long)
    for b, s in enumerate(seqs):
        out[b, : len(s)] = torch.tensor(s)
    return out
