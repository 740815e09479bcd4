"""Greedy generation for the functional models (reference experimental/core/generate/generate.py:9-110).  Model contract:
``forward(input_tokens, last_pos, attention_mask)`` -> logits ``[B, T, V]`` (the reference's) or next tokens ``[B]`` (models with on-device
arg-max).  Two call styles:

* ``generate(model, max_len, prompt_tokens: list[list[int]], stop_tokens, pad_token, return_logits=False) -> GenerateResult`` — the reference's;
* ``generate(model, prompt_tokens: Tensor, attention_mask=None, max_new_tokens=16, ...) -> Tensor`` — tensor in, padded tensor out."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn.functional as F

from .._legacy_generate import generate as _generate_tensor


@dataclass
class GenerateResult:
    prompt_tokens: List[List[int]]
    logits: Optional[List[torch.Tensor]] = None


def _next_tokens(out: torch.Tensor) -> torch.Tensor:
    return out[:, -1].argmax(-1) if out.is_floating_point() and out.dim() == 3 else out.reshape(out.shape[0], -1)[:, -1]


@torch.no_grad()
def _generate_lists(model, max_len: int, prompt_tokens: List[List[int]], stop_tokens: List[int], pad_token: int,
                    return_logits: bool = False) -> GenerateResult:
    prompts = [[t for t in p if t != pad_token] for p in prompt_tokens]
    B = len(prompts)
    last_pos = torch.tensor([len(p) - 1 for p in prompts], dtype=torch.int32)
    width = max(len(p) for p in prompts)
    tokens = torch.tensor([p + [pad_token] * (width - len(p)) for p in prompts], dtype=torch.int32)
    mask = torch.tensor([[1] * len(p) + [0] * (width - len(p)) for p in prompts], dtype=torch.int32)
    done = torch.zeros(B, dtype=torch.bool)
    logits_out = [] if return_logits else None
    if hasattr(model, "reset"):
        model.reset()
    inp = tokens
    while True:
        out = model.forward(input_tokens=inp, last_pos=last_pos, attention_mask=mask)
        if return_logits and out.is_floating_point() and out.dim() == 3:
            logits_out.append(out[:B, -1].float().cpu().clone())
        nxt = _next_tokens(out)[:B].cpu().to(torch.int32)
        last_pos = last_pos + 1
        for b in range(B):
            if not done[b]:
                prompts[b].append(int(nxt[b]))
        for s in stop_tokens:
            done |= nxt == s
        if bool(done.all()) or int(last_pos.max()) >= max_len:
            break
        inp = nxt.view(B, 1)
        mask = F.pad(mask, (0, 1), value=0)
        mask[torch.arange(B), last_pos.long()] = 1
    return GenerateResult(prompt_tokens=prompts, logits=logits_out)


def generate(model, *args, **kwargs):
    first = args[0] if args else kwargs.get("max_len", kwargs.get("prompt_tokens"))
    if isinstance(first, int) or isinstance(kwargs.get("prompt_tokens"), list):
        return _generate_lists(model, *args, **kwargs)
    return _generate_tensor(model, *args, **kwargs)
