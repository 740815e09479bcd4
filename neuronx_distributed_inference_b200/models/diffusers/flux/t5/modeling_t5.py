"""T5 encoder (sequence prompt embedding of FLUX; reference models/diffusers/flux/t5/modeling_t5.py).  Checkpoint layout:
transformers ``T5EncoderModel`` (v1.1: gated-GELU feed-forward, relative position bias owned by block 0, no attention scaling)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..... import ops
from .....modules.gqa import GroupQueryAttention_O, GroupQueryAttention_QKV
from .....modules.norm import RMSNorm
from .....modules.vision import VisionMLP
from .....parallel.state import get_tensor_model_parallel_group


def relative_position_bucket(rel, num_buckets=32, max_distance=128):
    num_buckets //= 2                                   # bidirectional
    ret = (rel > 0).long() * num_buckets
    n = rel.abs()
    max_exact = num_buckets // 2
    is_small = n < max_exact
    large = max_exact + (torch.log(n.float().clamp_min(1) / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = large.clamp(max=num_buckets - 1)
    return ret + torch.where(is_small, n, large)


class T5Block(nn.Module):
    def __init__(self, c, dtype, device):
        super().__init__()
        inner = c.num_heads * c.d_kv
        self.norm1 = RMSNorm(c.d_model, c.layer_norm_epsilon, dtype, device=device)
        self.norm2 = RMSNorm(c.d_model, c.layer_norm_epsilon, dtype, device=device)
        self.qkv = GroupQueryAttention_QKV(c.d_model, c.d_kv, c.num_heads, c.num_heads, None, dtype, False, None, device)
        self.o = GroupQueryAttention_O(c.d_model, c.d_kv, c.num_heads, c.num_heads, None, dtype, False, None, device)
        gated = "gated" in getattr(c, "feed_forward_proj", "gated-gelu")
        self.ff = VisionMLP(c.d_model, c.d_ff, "gelu_new" if gated else "relu", False, gated, dtype, device)
        self.h, self.d = self.qkv.n_q, c.d_kv

    def forward(self, x, bias, key_valid):
        B, N, _ = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(B, N, 3 * self.h, self.d).split(self.h, 2)
        s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) + bias       # T5: no 1/sqrt(d)
        if key_valid is not None:
            s = s.masked_fill(~key_valid.bool().view(B, 1, 1, N), float("-inf"))
        o = torch.einsum("bhqk,bkhd->bqhd", s.softmax(-1).to(v.dtype), v).reshape(B, N, self.h * self.d)
        x = self.o(o, x)
        return x + self.ff(self.norm2(x))


class NeuronT5EncoderModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        c, dt = config, config.neuron_config.torch_dtype
        self.config = config
        self.shared = nn.Embedding(c.vocab_size, c.d_model, dtype=dt, device=device)
        self.relative_attention_bias = nn.Embedding(c.relative_attention_num_buckets, c.num_heads, dtype=dt, device=device)
        self.blocks = nn.ModuleList([T5Block(c, dt, device) for _ in range(c.num_layers)])
        self.final_layer_norm = RMSNorm(c.d_model, c.layer_norm_epsilon, dt, device=device)
        g = get_tensor_model_parallel_group()
        self.tp_rank, self.tp = g.rank, g.size
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, input_ids, attention_mask=None):
        B, N = input_ids.shape
        x = self.shared(input_ids)
        pos = torch.arange(N, device=x.device)
        buckets = relative_position_bucket(pos[None, :] - pos[:, None], self.config.relative_attention_num_buckets,
                                           getattr(self.config, "relative_attention_max_distance", 128))
        bias = self.relative_attention_bias(buckets).permute(2, 0, 1).unsqueeze(0).float()       # [1,H,N,N]
        if self.tp > 1:
            hl = bias.shape[1] // self.tp
            bias = bias[:, self.tp_rank * hl:(self.tp_rank + 1) * hl]
        for blk in self.blocks:
            x = blk(x, bias, attention_mask)
        return self.final_layer_norm(x)


def convert_t5_state_dict(sd: dict, config) -> dict:
    out = {}
    for i in range(config.num_layers):
        a = f"encoder.block.{i}.layer.0.SelfAttention"
        out[f"blocks.{i}.qkv.weight"] = torch.cat([sd[f"{a}.q.weight"], sd[f"{a}.k.weight"], sd[f"{a}.v.weight"]], 0)
        out[f"blocks.{i}.o.weight"] = sd[f"{a}.o.weight"]
        out[f"blocks.{i}.norm1.weight"] = sd[f"encoder.block.{i}.layer.0.layer_norm.weight"]
        out[f"blocks.{i}.norm2.weight"] = sd[f"encoder.block.{i}.layer.1.layer_norm.weight"]
        f = f"encoder.block.{i}.layer.1.DenseReluDense"
        if f"{f}.wi_0.weight" in sd:
            out[f"blocks.{i}.ff.gate_up_proj.weight"] = torch.cat([sd[f"{f}.wi_0.weight"], sd[f"{f}.wi_1.weight"]], 0)
        else:
            out[f"blocks.{i}.ff.fc1.weight"] = sd[f"{f}.wi.weight"]
        out[f"blocks.{i}.ff.fc2.weight"] = sd[f"{f}.wo.weight"]
    out["relative_attention_bias.weight"] = sd["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    out["shared.weight"] = sd.get("shared.weight", sd.get("encoder.embed_tokens.weight"))
    out["final_layer_norm.weight"] = sd["encoder.final_layer_norm.weight"]
    return out
