"""HF -> engine state-dict conversions shared by the decoder models (per-model hooks of the
reference: modeling_llama.py:1207-1276, modeling_dbrx.py:51-112, modeling_qwen3_moe.py:121-235)."""
from __future__ import annotations

from typing import Dict

import torch


def _cat(ts):
    if ts[0].dtype in (torch.float8_e4m3fn, torch.float8_e5m2):
        return torch.cat([t.view(torch.uint8) for t in ts], 0).view(ts[0].dtype)
    return torch.cat(ts, 0)


def fuse_qkv_and_gate_up(sd: Dict[str, torch.Tensor], num_layers: int, prefix: str = "layers.",
                         attn: str = "self_attn", mlp: str = "mlp", fuse_mlp: bool = True) -> Dict[str, torch.Tensor]:
    """q/k/v -> ``qkv_proj`` ([q;k;v] on dim 0) and gate/up -> ``gate_up_proj`` for weights, biases
    and per-channel quantisation scales alike."""
    sd = dict(sd)
    for i in range(num_layers):
        a = f"{prefix}{i}.{attn}."
        for suffix in ("weight", "bias", "scale"):
            ks = [f"{a}{p}_proj.{suffix}" for p in "qkv"]
            if all(k in sd for k in ks):
                parts = [sd.pop(k) for k in ks]
                if suffix == "scale" and all(p.numel() == 1 for p in parts):
                    # per-tensor scales cannot be concatenated: expand to per-channel
                    w = sd.get(f"{a}qkv_proj.weight")
                    raise ValueError("per-tensor scales must be fused before weights are quantised "
                                     f"({a}qkv_proj)") if w is None else None
                sd[f"{a}qkv_proj.{suffix}"] = _cat(parts)
        if not fuse_mlp:
            continue
        m = f"{prefix}{i}.{mlp}."
        for suffix in ("weight", "bias", "scale"):
            ks = [f"{m}gate_proj.{suffix}", f"{m}up_proj.{suffix}"]
            if all(k in sd for k in ks):
                sd[f"{m}gate_up_proj.{suffix}"] = _cat([sd.pop(k) for k in ks])
    return {k: v for k, v in sd.items() if "rotary_emb.inv_freq" not in k}
