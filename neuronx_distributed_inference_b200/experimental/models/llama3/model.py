"""Functional Llama-3 (reference experimental/models/llama3/model.py:1-673): a plain module whose forward is
``forward(input_tokens, last_pos, attention_mask)`` built only from :mod:`experimental.functional` calls on explicit weight
tensors — no application/wrapper layers; useful as the smallest end-to-end example of the kernels."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn as nn

from .... import ops
from ....modules.rope import RotaryEmbedding
from ... import functional as F


@dataclass
class Llama3Args:
    dim: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    n_kv_heads: int = 8
    vocab_size: int = 128256
    ffn_dim: int = 14336
    norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = None
    max_batch_size: int = 2
    max_seq_len: int = 2048
    dtype: torch.dtype = torch.bfloat16


class Llama3(nn.Module):
    def __init__(self, args: Llama3Args, weights: Dict[str, torch.Tensor], device=None):
        """``weights`` uses the engine's converted names (``layers.i.self_attn.qkv_proj.weight`` ...)."""
        super().__init__()
        self.a = a = args
        self.hd = a.dim // a.n_heads
        self.w = {k: v.to(device=device, dtype=a.dtype if v.is_floating_point() else v.dtype) for k, v in weights.items()}
        self.rope = RotaryEmbedding(self.hd, a.max_seq_len, a.rope_theta, a.rope_scaling, device=device)
        self.k = torch.zeros(a.n_layers, a.max_batch_size + 1, a.n_kv_heads, a.max_seq_len, self.hd, dtype=a.dtype, device=device)
        self.v = torch.zeros_like(self.k)

    def reset(self):
        self.k.zero_()
        self.v.zero_()

    @torch.no_grad()
    def forward(self, input_tokens: torch.Tensor, last_pos: torch.Tensor, attention_mask: Optional[torch.Tensor] = None):
        """Prefill when ``input_tokens`` has more than one column (positions 0..T-1, ``last_pos`` = index of each row's last
        real token), decode otherwise (``last_pos`` = position of the token).  -> next token ids [B]."""
        a, w, hd = self.a, self.w, self.hd
        B, T = input_tokens.shape
        dev = self.k.device
        tok = input_tokens.to(dev)
        last_pos = last_pos.to(dev)
        prefill = T > 1
        pos = torch.arange(T, device=dev).view(1, T).expand(B, T) if prefill else last_pos.view(B, 1)
        write = pos if attention_mask is None or not prefill else torch.where(attention_mask.to(dev).bool(), pos, torch.full_like(pos, -1))
        seq = torch.arange(B, device=dev, dtype=torch.int32)
        cos, sin = self.rope(pos)
        h = nn.functional.embedding(tok, w["embed_tokens.weight"])
        for i in range(a.n_layers):
            p = f"layers.{i}."
            q, k, v = F.qkv_proj(h, w[p + "self_attn.qkv_proj.weight"], a.n_heads, a.n_kv_heads, hd,
                                 norm_weight=w[p + "input_layernorm.weight"], norm_eps=a.norm_eps)
            q, k = ops.apply_rope(q, cos, sin, False), ops.apply_rope(k, cos, sin, False)
            ops.kv_append(self.k[i], self.v[i], k, v, seq, write.to(torch.int32))
            if prefill:
                o = F.causal_scaled_dot_product_attention(q, k, v)
            else:
                o = ops.attention_decode(q, self.k[i], self.v[i], seq, pos.to(torch.int32), hd ** -0.5)
            h = F.o_proj_allreduce(o.reshape(B, T, -1), w[p + "self_attn.o_proj.weight"], residual=h)
            h = F.gated_mlp_fused(h, w[p + "mlp.gate_up_proj.weight"], w[p + "mlp.down_proj.weight"],
                                  norm_weight=w[p + "post_attention_layernorm.weight"], norm_eps=a.norm_eps, residual=h)
        hl = h[torch.arange(B, device=dev), last_pos.long()] if prefill else h[:, 0]
        logits = ops.linear(hl.unsqueeze(1), w["lm_head.weight"], None, norm_weight=w["norm.weight"], norm_eps=a.norm_eps)
        return ops.argmax(logits[:, 0].float())
