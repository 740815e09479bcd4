"""Functional Llama-4 text decoder (reference experimental/models/llama4/model.py:1-724): like the functional Llama-3, one module
whose forward is written directly in :mod:`experimental.functional` / ``ops`` calls on explicit weight tensors.

Per layer: interleaved RoPE + chunked attention + L2 q/k norm on the "local" layers, NO positions + global attention + query
temperature scaling on every 4th ("NoPE") layer; dense SwiGLU or a top-k sigmoid-routed MoE whose affinity scales the expert INPUT,
plus an always-on shared expert."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .... import ops
from ....modules.rope import RotaryEmbedding
from ... import functional as F
from ..config import Config


class Llama4(nn.Module):
    def __init__(self, cfg: Config, weights: Dict[str, torch.Tensor], device=None):
        """``weights``: the engine's converted names (``layers.i.mlp.expert_mlps.gate_up_proj`` [E,2I,H] ...)."""
        super().__init__()
        self.c = c = cfg
        self.w = {k: v.to(device=device, dtype=c.dtype if (v.is_floating_point() and "router" not in k) else v.dtype) for k, v in weights.items()}
        self.rope = RotaryEmbedding(c.head_dim, c.max_seq_len, c.rope_theta, c.rope_scaling, device=device)
        self.k = torch.zeros(c.n_layers, c.max_batch_size + 1, c.n_kv_heads, c.max_seq_len, c.head_dim, dtype=c.dtype, device=device)
        self.v = torch.zeros_like(self.k)
        self.nope = set(c.nope_layers or [])
        self.moe = set(c.moe_layers or [])

    def reset(self):
        self.k.zero_()
        self.v.zero_()

    def _l2(self, x):
        return (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + self.c.norm_eps)).to(x.dtype)

    def _moe(self, x, p):
        c, w = self.c, self.w
        B, T, H = x.shape
        x2 = x.reshape(-1, H)
        logits = ops.ref.linear(x2.float(), w[p + "mlp.router.linear_router.weight"].float())
        top, idx = logits.topk(c.top_k, -1)
        aff = torch.sigmoid(top)                                           # sigmoid AFTER the top-k selection
        y = ops.moe_experts(x2, w[p + "mlp.expert_mlps.gate_up_proj"], w[p + "mlp.expert_mlps.down_proj"], aff, idx, scale_input=True)
        y = y + F.gated_mlp_kernel_unreduced(x2.unsqueeze(0), w[p + "mlp.shared_experts.gate_up_proj.weight"],
                                             w[p + "mlp.shared_experts.down_proj.weight"]).squeeze(0)
        return y.view(B, T, H)

    @torch.no_grad()
    def forward(self, input_tokens: torch.Tensor, last_pos: torch.Tensor, attention_mask: Optional[torch.Tensor] = None):
        """Same calling convention as the functional Llama-3.  -> next token ids [B]."""
        c, w, hd = self.c, self.w, self.c.head_dim
        B, T = input_tokens.shape
        dev = self.k.device
        tok, last_pos = input_tokens.to(dev), last_pos.to(dev)
        prefill = T > 1
        pos = torch.arange(T, device=dev).view(1, T).expand(B, T) if prefill else last_pos.view(B, 1)
        write = pos if attention_mask is None or not prefill else torch.where(attention_mask.to(dev).bool(), pos, torch.full_like(pos, -1))
        seq = torch.arange(B, device=dev, dtype=torch.int32)
        cos, sin = self.rope(pos)
        h = nn.functional.embedding(tok, w["embed_tokens.weight"])
        for i in range(c.n_layers):
            p = f"layers.{i}."
            local = i not in self.nope
            q, k, v = F.qkv_proj(h, w[p + "self_attn.qkv_proj.weight"], c.n_heads, c.n_kv_heads, hd,
                                 norm_weight=w[p + "input_layernorm.weight"], norm_eps=c.norm_eps)
            if local:
                q, k = ops.apply_rope(q, cos, sin, True), ops.apply_rope(k, cos, sin, True)
                if c.use_qk_norm:
                    q, k = self._l2(q), self._l2(k)
            elif c.attn_temperature_tuning:
                s = torch.log1p(torch.floor((pos.float() + 1.0) / c.floor_scale)) * c.attn_scale + 1.0
                q = (q * s.view(B, T, 1, 1)).to(q.dtype)
            chunk = c.attention_chunk_size if local else None
            ops.kv_append(self.k[i], self.v[i], k, v, seq, write.to(torch.int32))
            if prefill:
                o = ops.attention_prefill(q, k, v, hd ** -0.5, True, None, chunk)
            else:
                o = ops.attention_decode(q, self.k[i], self.v[i], seq, pos.to(torch.int32), hd ** -0.5, None, chunk)
            h = F.o_proj_allreduce(o.reshape(B, T, -1), w[p + "self_attn.o_proj.weight"], residual=h)
            if i in self.moe:
                h = h + self._moe(ops.rmsnorm(h, w[p + "post_attention_layernorm.weight"], c.norm_eps), p)
            else:
                h = F.gated_mlp_fused(h, w[p + "mlp.gate_up_proj.weight"], w[p + "mlp.down_proj.weight"],
                                      norm_weight=w[p + "post_attention_layernorm.weight"], norm_eps=c.norm_eps, residual=h)
        hl = h[torch.arange(B, device=dev), last_pos.long()] if prefill else h[:, 0]
        logits = ops.linear(hl.unsqueeze(1), w["lm_head.weight"], None, norm_weight=w["norm.weight"], norm_eps=c.norm_eps)
        return ops.argmax(logits[:, 0].float())
