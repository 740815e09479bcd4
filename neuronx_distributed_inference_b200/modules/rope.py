"""Rotary embeddings: default, Llama-3 frequency scaling, linear / dynamic-NTK, YaRN (GPT-OSS /
DeepSeek), partial rotary, M-RoPE sections (Qwen2-VL).  The tables are built once in fp32 as
``cos/sin [max_pos, rot_dim/2]`` device tensors; kernels gather rows by position.
reference: modules/attention/utils.py:200-249 (RotaryEmbedding), models/llama/modeling_llama.py:805-870
(Llama3RotaryEmbedding), models/deepseek/rope_util.py.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn


def _yarn_find_dim(num_rot, dim, base, max_pos):
    return (dim * math.log(max_pos / (num_rot * 2 * math.pi))) / (2 * math.log(base))


def compute_inv_freq(dim: int, base: float, scaling: Optional[dict] = None, max_pos: int = 8192):
    """-> (inv_freq [dim/2] fp32, attention scale factor applied to cos/sin)."""
    inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    mscale = 1.0
    if not scaling:
        return inv, mscale
    kind = scaling.get("rope_type", scaling.get("type", "default"))
    factor = float(scaling.get("factor", 1.0))
    if kind in ("default", None):
        pass
    elif kind == "linear":
        inv = inv / factor
    elif kind == "llama3":
        lo = scaling.get("low_freq_factor", 1.0)
        hi = scaling.get("high_freq_factor", 4.0)
        old = scaling.get("original_max_position_embeddings", 8192)
        wavelen = 2 * math.pi / inv
        smooth = ((old / wavelen) - lo) / (hi - lo)
        scaled = torch.where(wavelen > old / lo, inv / factor, inv)
        mid = (wavelen <= old / lo) & (wavelen >= old / hi)
        inv = torch.where(mid, (1 - smooth) * inv / factor + smooth * inv, scaled)
    elif kind == "dynamic":
        old = scaling.get("original_max_position_embeddings", max_pos)
        if max_pos > old:
            base = base * ((factor * max_pos / old) - (factor - 1)) ** (dim / (dim - 2))
            inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    elif kind == "yarn":
        old = scaling.get("original_max_position_embeddings", max_pos)
        beta_fast = scaling.get("beta_fast", 32)
        beta_slow = scaling.get("beta_slow", 1)
        lo = max(math.floor(_yarn_find_dim(beta_fast, dim, base, old)), 0)
        hi = min(math.ceil(_yarn_find_dim(beta_slow, dim, base, old)), dim - 1)
        if scaling.get("truncate", True) is False:
            lo = max(_yarn_find_dim(beta_fast, dim, base, old), 0)
            hi = min(_yarn_find_dim(beta_slow, dim, base, old), dim - 1)
        ramp = ((torch.arange(dim // 2, dtype=torch.float32) - lo) / max(hi - lo, 1e-3)).clamp(0, 1)
        extrap_mask = 1 - ramp
        inv = inv / factor * (1 - extrap_mask) + inv * extrap_mask
        ms = scaling.get("mscale", 1.0)
        msa = scaling.get("mscale_all_dim", 0.0)

        def get_mscale(s, m=1.0):
            return 1.0 if s <= 1 else 0.1 * m * math.log(s) + 1.0
        if "attention_factor" in scaling and scaling["attention_factor"] is not None:
            mscale = float(scaling["attention_factor"])
        elif msa:
            mscale = get_mscale(factor, ms) / get_mscale(factor, msa)
        else:
            mscale = get_mscale(factor)
    else:
        raise ValueError(f"unsupported rope scaling {kind}")
    return inv, mscale


class RotaryEmbedding(nn.Module):
    """Holds cos/sin tables; ``forward(position_ids [B,T])`` -> (cos, sin) [B,T,rot_dim/2] fp32."""

    def __init__(self, dim: int, max_position_embeddings: int = 2048, base: float = 10000.0,
                 scaling: Optional[dict] = None, device=None):
        super().__init__()
        self.dim = dim
        self.max_position_embeddings = max_position_embeddings
        self.base = base
        inv, self.mscale = compute_inv_freq(dim, base, scaling, max_position_embeddings)
        self.register_buffer("inv_freq", inv.to(device) if device is not None else inv, persistent=False)
        self._table_len = 0
        self.register_buffer("cos_table", torch.empty(0), persistent=False)
        self.register_buffer("sin_table", torch.empty(0), persistent=False)

    def build_tables(self, length: int, device=None):
        device = device or self.inv_freq.device
        t = torch.arange(length, dtype=torch.float32, device=device)
        fr = torch.outer(t, self.inv_freq.to(device))
        self.cos_table = (fr.cos() * self.mscale).contiguous()
        self.sin_table = (fr.sin() * self.mscale).contiguous()
        self._table_len = length

    def forward(self, position_ids: torch.Tensor):
        need = self.max_position_embeddings
        if self._table_len < need or self.cos_table.device != position_ids.device:
            self.build_tables(need, position_ids.device)
        p = position_ids.long().clamp(0, self._table_len - 1)
        return self.cos_table[p], self.sin_table[p]


class MRotaryEmbedding(RotaryEmbedding):
    """Multimodal RoPE (Qwen2-VL): position_ids [3,B,T] (t,h,w), frequency bands split by
    ``mrope_section``; falls back to ordinary RoPE for 2-D position ids."""

    def __init__(self, dim, max_position_embeddings, base, mrope_section, scaling=None, device=None, interleaved: bool = False):
        super().__init__(dim, max_position_embeddings, base, None, device)
        self.mrope_section = list(mrope_section)
        self.interleaved = interleaved

    def forward(self, position_ids):
        if position_ids.dim() == 2:
            return super().forward(position_ids)
        cs = [super(MRotaryEmbedding, self).forward(position_ids[i]) for i in range(3)]
        if self.interleaved:
            # Qwen3-VL: bands interleaved [T H W T H W ... T T] instead of chunked [T.. H.. W..]
            cos, sin = cs[0][0].clone(), cs[0][1].clone()
            for axis, off in ((1, 1), (2, 2)):
                idx = slice(off, self.mrope_section[axis] * 3, 3)
                cos[..., idx] = cs[axis][0][..., idx]
                sin[..., idx] = cs[axis][1][..., idx]
            return cos, sin
        cos_parts, sin_parts, o = [], [], 0
        for i, sec in enumerate(self.mrope_section):
            cos_parts.append(cs[i % 3][0][..., o:o + sec])
            sin_parts.append(cs[i % 3][1][..., o:o + sec])
            o += sec
        return torch.cat(cos_parts, -1), torch.cat(sin_parts, -1)
