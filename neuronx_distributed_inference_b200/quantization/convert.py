"""Weight quantisation: offline checkpoint generation + module conversion.

Rebuild of the external ``neuronx_distributed.quantization`` surface the reference drives:
``quantize_pytorch_model_per_{tensor,channel}_symmetric`` + ``prepare_quantized_state_dict`` (application_base.py:746-799)
and ``convert(model, q_config, modules_to_not_convert)`` swapping parallel linears for quantised ones holding
``weight`` (int8 / fp8) + ``scale`` (fp32, never down-cast; model_wrapper.py:1477-1529).

B200 mapping: weights stay 8-bit in HBM; the decode GEMV dequantises in registers (half the bytes of bf16 => ~2x the
decode roofline), the prefill GEMM dequantises on the fly as well (``ops.linear(..., scale=...)``).  Per-channel scales
are per *output* channel, so they shard with the weight rows (Column) or replicate (Row) — the GQA index plan shards
scales with the same gather as weights (modules/gqa.py).
"""
from __future__ import annotations

import fnmatch
from typing import Dict, Iterable, List, Optional

import torch
import torch.nn as nn

from ..ops import reference as ref

_DT = {"int8": torch.int8, "f8e4m3": torch.float8_e4m3fn, "f8e5m2": torch.float8_e5m2}


def _skip(name: str, modules_to_not_convert: Optional[Iterable[str]]) -> bool:
    for pat in modules_to_not_convert or ():
        if pat in name or fnmatch.fnmatch(name, pat) or fnmatch.fnmatch(name, pat + ".*"):
            return True
    return False


def _is_linear_weight(k: str, v: torch.Tensor) -> bool:
    return k.endswith(".weight") and v.dim() == 2 and not any(s in k for s in ("embed_tokens", "norm", "router", "lm_head_norm"))


_EXPERT_KEYS = (".expert_mlps.gate_up_proj", ".expert_mlps.down_proj")


def quantize_experts(w: torch.Tensor, dtype):
    """``[E, out, in]`` expert bank -> (8-bit bank, fp32 scales ``[E, out]``): symmetric, one scale per expert AND output channel
    (reference ``expert_wise_per_channel_symmetric``)."""
    qmax = 127.0 if dtype == torch.int8 else float(torch.finfo(dtype).max)
    s = (w.float().abs().amax(-1) / qmax).clamp_min(1e-12)
    q = w.float() / s.unsqueeze(-1)
    q = q.round().clamp(-qmax, qmax).to(torch.int8) if dtype == torch.int8 else q.clamp(-qmax, qmax).to(dtype)
    return q, s


def quantize_state_dict(sd: Dict[str, torch.Tensor], neuron_config, is_draft: bool = False) -> Dict[str, torch.Tensor]:
    """Full-precision converted state dict -> quantised state dict (``X.weight`` 8-bit + ``X.scale`` fp32; MoE expert banks
    ``X.expert_mlps.{gate_up,down}_proj`` 8-bit + ``..._scale`` when the type is expert-wise per channel)."""
    qt, qd = neuron_config.quantization_type, _DT[neuron_config.quantization_dtype]
    skip = neuron_config.draft_model_modules_to_not_convert if is_draft else neuron_config.modules_to_not_convert
    out = {}
    for k, v in sd.items():
        if qt == "expert_wise_per_channel_symmetric" and k.endswith(_EXPERT_KEYS) and v.dim() == 3 and v.is_floating_point() \
                and not _skip(k, skip):
            out[k], out[k.replace("_proj", "_scale")] = quantize_experts(v, qd)
            continue
        mod = k[: -len(".weight")] if k.endswith(".weight") else k
        if not _is_linear_weight(k, v) or _skip(mod, skip) or not v.is_floating_point():
            out[k] = v
            continue
        if qt == "per_tensor_symmetric":
            q, s = ref.quantize_per_tensor(v, qd)
        elif qt in ("per_channel_symmetric", "expert_wise_per_channel_symmetric"):
            q, s = ref.quantize_per_channel(v, qd)
        elif qt == "blockwise_symmetric":
            bs = neuron_config.quantization_block_size or [128, 128]
            q, s = ref.quantize_blockwise(v, tuple(bs), qd)
        else:
            raise ValueError(qt)
        out[k] = q
        out[mod + ".scale"] = s.float()
    return out


def prepare_quantized_state_dict(sd):
    """Kept for API parity (the reference converts torch ``qint8`` to ``int8`` here); ours is already plain."""
    return {k: v for k, v in sd.items() if v is not None}


def convert(model: nn.Module, neuron_config, modules_to_not_convert: Optional[List[str]] = None, is_draft: bool = False):
    """In place: every quantisable parallel linear gets an 8-bit ``weight`` and an fp32 ``scale`` parameter with the same
    sharding metadata; its forward then passes ``scale`` to ``ops.linear`` / ``ops.linear_allreduce``."""
    from ..parallel.layers import BaseParallelLinear
    qd = _DT[neuron_config.quantization_dtype]
    per_tensor = neuron_config.quantization_type == "per_tensor_symmetric"
    blockwise = neuron_config.quantization_type == "blockwise_symmetric"
    skip = modules_to_not_convert
    if skip is None:
        skip = neuron_config.draft_model_modules_to_not_convert if is_draft else neuron_config.modules_to_not_convert
    n = 0
    if neuron_config.quantization_type == "expert_wise_per_channel_symmetric":
        from ..modules.moe import ExpertMLPs
        for name, mod in model.named_modules():
            if isinstance(mod, ExpertMLPs) and not _skip(name, skip):
                _convert_experts(mod, qd)
                n += 1
    for name, mod in model.named_modules():
        if not isinstance(mod, BaseParallelLinear) or not getattr(mod, "quantizable", True) or _skip(name, skip):
            continue
        if name.endswith("lm_head") and not any("lm_head" in s for s in (skip or [])) is False:
            continue
        w = mod.weight
        new_w = nn.Parameter(torch.zeros(w.shape, dtype=torch.int8, device=w.device).view(qd) if qd != torch.int8
                             else torch.zeros(w.shape, dtype=torch.int8, device=w.device), requires_grad=False)
        for a in ("partition_dim", "tp_group", "partition_stride", "shard_fn"):
            if hasattr(w, a):
                setattr(new_w, a, getattr(w, a))
        mod.weight = new_w
        if per_tensor:
            sc = nn.Parameter(torch.ones(1, dtype=torch.float32, device=w.device), requires_grad=False)
            sc.partition_dim, sc.tp_group = None, getattr(w, "tp_group", None)
        elif blockwise:
            bs = neuron_config.quantization_block_size or [128, 128]
            sc = nn.Parameter(torch.ones(-(-w.shape[0] // bs[0]), -(-w.shape[1] // bs[1]), dtype=torch.float32, device=w.device),
                              requires_grad=False)
            sc.partition_dim, sc.tp_group = getattr(w, "partition_dim", None), getattr(w, "tp_group", None)
            sc.partition_stride = getattr(w, "partition_stride", 1)
        else:
            sc = nn.Parameter(torch.ones(w.shape[0], dtype=torch.float32, device=w.device), requires_grad=False)
            pdim = getattr(w, "partition_dim", None)
            # per-output-channel scale follows the weight only when the OUTPUT dim is the sharded one
            sc.partition_dim = 0 if pdim == 0 else None
            sc.tp_group = getattr(w, "tp_group", None)
            sc.partition_stride = getattr(w, "partition_stride", 1)
            if hasattr(w, "shard_fn"):
                sc.shard_fn = w.shard_fn
        mod.scale = sc
        n += 1
    return n


def _convert_experts(mod, qd):
    """8-bit expert banks + fp32 scales ``[E_local, out_local]`` sharded like the rows of their weights."""
    tps, tpr = mod.tp_group.size, mod.tp_group.rank
    E0, El = mod.expert_offset, mod.E_local
    for wname, sname in (("gate_up_proj", "gate_up_scale"), ("down_proj", "down_scale")):
        w = getattr(mod, wname)
        new_w = nn.Parameter(torch.zeros(w.shape, dtype=torch.int8, device=w.device).view(qd) if qd != torch.int8
                             else torch.zeros(w.shape, dtype=torch.int8, device=w.device), requires_grad=False)
        for a in ("partition_dim", "tp_group", "shard_fn"):
            if hasattr(w, a):
                setattr(new_w, a, getattr(w, a))
        setattr(mod, wname, new_w)
        sc = nn.Parameter(torch.ones(w.shape[0], w.shape[1], dtype=torch.float32, device=w.device), requires_grad=False)
        sc.partition_dim, sc.tp_group = 0, getattr(w, "tp_group", None)
        if wname == "gate_up_proj":      # [E, 2I]: experts sliced, gate and up halves sharded over the MoE-TP ranks like the weight rows
            sc.shard_fn = lambda full, rank: torch.cat([h.chunk(tps, 1)[tpr] for h in full[E0:E0 + El].chunk(2, 1)], 1).contiguous()
        else:                            # [E, H]: the output dim of down_proj is not sharded
            sc.shard_fn = lambda full, rank: full[E0:E0 + El].contiguous()
        setattr(mod, sname, sc)
