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


def convert_moe_experts(sd: Dict[str, torch.Tensor], num_layers: int, num_experts: int, moe_prefixes=("mlp",),
                        gate_names=("gate",), w_names=("w1", "w3", "w2"), layers=None, prefix: str = "layers.",
                        dst: str = "mlp") -> Dict[str, torch.Tensor]:
    """HF MoE block -> engine names:  ``<dst>.router.linear_router.weight`` [E,H] fp32,
    ``<dst>.expert_mlps.gate_up_proj`` [E,2I,H], ``<dst>.expert_mlps.down_proj`` [E,H,I].
    Accepts the transformers>=5 fused tensors (``experts.gate_up_proj`` [E,2I,H] / ``experts.down_proj`` [E,H,I]) and
    the per-expert hub format (``experts.{e}.<gate|up|down>.weight``)."""
    sd = dict(sd)
    g_name, u_name, d_name = w_names
    for i in (layers if layers is not None else range(num_layers)):
        for mp in moe_prefixes:
            base = f"{prefix}{i}.{mp}."
            tgt = f"{prefix}{i}.{dst}."
            for gn in gate_names:
                if base + gn + ".weight" in sd:
                    sd[tgt + "router.linear_router.weight"] = sd.pop(base + gn + ".weight").float()
                if base + gn + ".bias" in sd:
                    sd[tgt + "router.linear_router.bias"] = sd.pop(base + gn + ".bias").float()
            if base + "experts.gate_up_proj" in sd:
                sd[tgt + "expert_mlps.gate_up_proj"] = sd.pop(base + "experts.gate_up_proj")
                sd[tgt + "expert_mlps.down_proj"] = sd.pop(base + "experts.down_proj")
            elif f"{base}experts.0.{g_name}.weight" in sd:
                gs = [sd.pop(f"{base}experts.{e}.{g_name}.weight") for e in range(num_experts)]
                us = [sd.pop(f"{base}experts.{e}.{u_name}.weight") for e in range(num_experts)]
                ds = [sd.pop(f"{base}experts.{e}.{d_name}.weight") for e in range(num_experts)]
                sd[tgt + "expert_mlps.gate_up_proj"] = torch.stack([torch.cat([g, u], 0) for g, u in zip(gs, us)])
                sd[tgt + "expert_mlps.down_proj"] = torch.stack(ds)
            elif f"{base}experts.0.gate_up_proj.weight" in sd:   # fused-by-the-generic-pass per-expert tensors
                sd[tgt + "expert_mlps.gate_up_proj"] = torch.stack(
                    [sd.pop(f"{base}experts.{e}.gate_up_proj.weight") for e in range(num_experts)])
                sd[tgt + "expert_mlps.down_proj"] = torch.stack(
                    [sd.pop(f"{base}experts.{e}.down_proj.weight") for e in range(num_experts)])
    return sd


def dequantize_block_fp8(sd: Dict[str, torch.Tensor], dtype=torch.bfloat16, block=(128, 128)) -> Dict[str, torch.Tensor]:
    """HF block-fp8 checkpoints (``*.weight`` e4m3 + ``*.weight_scale_inv`` [out/128, in/128]) -> dense ``dtype``
    (reference qwen3_moe ``maybe_dequantize_layer`` :103-118)."""
    out = {}
    for k, v in sd.items():
        if k.endswith("_scale_inv") or (k.endswith(".scale") and k[:-6] + ".weight" in sd and sd[k[:-6] + ".weight"].dtype == torch.float8_e4m3fn and v.dim() == 2):
            continue
        s = sd.get(k + "_scale_inv")
        if s is not None and v.dtype in (torch.float8_e4m3fn, torch.float8_e5m2):
            w = v.float()
            r0 = -(-w.shape[0] // s.shape[0])
            r1 = -(-w.shape[1] // s.shape[1])
            sc = s.float().repeat_interleave(r0, 0)[: w.shape[0]].repeat_interleave(r1, 1)[:, : w.shape[1]]
            out[k] = (w * sc).to(dtype)
        else:
            out[k] = v
    return out
