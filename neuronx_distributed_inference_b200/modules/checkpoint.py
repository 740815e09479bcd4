"""Checkpoint I/O and tensor-parallel sharding.

reference: modules/checkpoint.py:24-285 (HF safetensors / sharded safetensors / .bin loading, sharded
safetensors saving, ``create_n_layer_checkpoint``, ``prune_state_dict``) and the external
``ModelBuilder.shard_checkpoint`` walk.  Sharding here is driven by per-parameter
metadata set by the parallel layers (``partition_dim``, ``partition_stride``, optional ``shard_fn``).
"""
from __future__ import annotations

import glob
import json
import os
import re
from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn

SAFETENSORS = "model.safetensors"
SAFETENSORS_INDEX = "model.safetensors.index.json"
PT_BIN = "pytorch_model.bin"
PT_BIN_INDEX = "pytorch_model.bin.index.json"
_DIFFUSERS = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.safetensors.index.json")


def _load_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """Load a full (unsharded) state dict from a directory or a single file."""
    if os.path.isfile(path):
        return _load_file(path)
    if not os.path.isdir(path):
        raise FileNotFoundError(f"{path} is neither a file nor a directory")
    for single in (SAFETENSORS, _DIFFUSERS[0], PT_BIN):
        p = os.path.join(path, single)
        if os.path.isfile(p):
            return _load_file(p)
    for index in (SAFETENSORS_INDEX, _DIFFUSERS[1], PT_BIN_INDEX):
        p = os.path.join(path, index)
        if os.path.isfile(p):
            with open(p) as f:
                files = sorted(set(json.load(f)["weight_map"].values()))
            sd = {}
            for fn in files:
                sd.update(_load_file(os.path.join(path, fn)))
            return sd
    files = sorted(glob.glob(os.path.join(path, "*.safetensors"))) or sorted(glob.glob(os.path.join(path, "*.pt")))
    if files:
        sd = {}
        for fn in files:
            sd.update(_load_file(fn))
        return sd
    raise FileNotFoundError(f"no checkpoint files found under {path}")


def save_state_dict_safetensors(state_dict: Dict[str, torch.Tensor], out_dir: str, max_shard_size: int = 5 << 30):
    """Write (sharded) safetensors + index, HF-compatible."""
    from safetensors.torch import save_file
    os.makedirs(out_dir, exist_ok=True)
    shards, cur, size = [], {}, 0
    for k, v in state_dict.items():
        v = v.detach().contiguous().cpu()
        nbytes = v.numel() * v.element_size()
        if cur and size + nbytes > max_shard_size:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = v
        size += nbytes
    if cur:
        shards.append(cur)
    if len(shards) == 1:
        save_file(shards[0], os.path.join(out_dir, SAFETENSORS), metadata={"format": "pt"})
        return
    weight_map, total = {}, 0
    for i, sh in enumerate(shards):
        fn = f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, os.path.join(out_dir, fn), metadata={"format": "pt"})
        for k, v in sh.items():
            weight_map[k] = fn
            total += v.numel() * v.element_size()
    with open(os.path.join(out_dir, SAFETENSORS_INDEX), "w") as f:
        json.dump({"metadata": {"total_size": total}, "weight_map": weight_map}, f, indent=2)


def prune_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k: v for k, v in sd.items() if v is not None}


def create_n_layer_checkpoint(src_dir: str, dst_dir: str, n_layers: int, layer_key: str = r"layers\.(\d+)\."):
    """Keep the first ``n_layers`` decoder layers (tiny test checkpoints; reference checkpoint.py:202-285)."""
    sd = load_state_dict(src_dir)
    pat = re.compile(layer_key)
    out = {}
    for k, v in sd.items():
        m = pat.search(k)
        if m and int(m.group(1)) >= n_layers:
            continue
        out[k] = v
    save_state_dict_safetensors(out, dst_dir)
    cfg = os.path.join(src_dir, "config.json")
    if os.path.isfile(cfg):
        with open(cfg) as f:
            c = json.load(f)
        for key in ("num_hidden_layers", "n_layers", "num_layers"):
            if key in c:
                c[key] = n_layers
        with open(os.path.join(dst_dir, "config.json"), "w") as f:
            json.dump(c, f, indent=2)
    return out


# ---- TP sharding ------------------------------------------------------------------------------
def shard_tensor(full: torch.Tensor, param: nn.Parameter, rank: int, size: int) -> torch.Tensor:
    fn = getattr(param, "shard_fn", None)
    if fn is not None:
        return fn(full, rank)
    dim = getattr(param, "partition_dim", None)
    if dim is None or size == 1:
        return full
    if full.dim() <= dim:  # e.g. per-tensor scale
        return full
    stride = getattr(param, "partition_stride", 1)
    if full.shape[dim] % (size * stride) != 0:
        pad = (-full.shape[dim]) % (size * stride)
        shape = list(full.shape)
        shape[dim] = pad
        full = torch.cat([full, full.new_zeros(shape)], dim)
    if stride == 1:
        return full.chunk(size, dim)[rank]
    blocks = full.chunk(stride, dim)
    return torch.cat([b.chunk(size, dim)[rank] for b in blocks], dim)


def shard_state_dict(model: nn.Module, full_sd: Dict[str, torch.Tensor], strict: bool = True,
                     rank_override: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Slice a full state dict down to what this rank's ``model`` parameters expect."""
    out, missing = {}, []
    for name, p in list(model.named_parameters()) + list(model.named_buffers()):
        if name not in full_sd:
            if isinstance(p, nn.Parameter) or name in dict(model.named_buffers()) and _persistent(model, name):
                missing.append(name)
            continue
        g = getattr(p, "tp_group", None)
        rank = g.rank if g is not None else 0
        size = g.size if g is not None else 1
        if rank_override is not None and g is not None:
            rank = rank_override
        t = shard_tensor(full_sd[name], p, rank, size)
        if tuple(t.shape) != tuple(p.shape):
            raise ValueError(f"{name}: sharded checkpoint tensor {tuple(t.shape)} != parameter {tuple(p.shape)}")
        out[name] = t
    if strict and missing:
        raise KeyError(f"missing keys in checkpoint: {missing[:8]}{' ...' if len(missing) > 8 else ''}")
    model._load_report = {"missing": missing, "unexpected": sorted(set(full_sd) - set(out))}
    return out


def _persistent(model, name):
    mod_name, _, leaf = name.rpartition(".")
    mod = model.get_submodule(mod_name) if mod_name else model
    return leaf not in getattr(mod, "_non_persistent_buffers_set", set())


def load_sharded(model: nn.Module, full_sd: Dict[str, torch.Tensor], dtype: Optional[torch.dtype] = None,
                 strict: bool = True):
    """Shard + cast + copy into the model's (possibly CUDA) parameters.  Floating tensors are cast
    to ``dtype`` except fp8 weights and ``*.scale`` (reference application_base.py:645-681)."""
    sd = shard_state_dict(model, full_sd, strict)
    params = dict(model.named_parameters())
    params.update(dict(model.named_buffers()))
    with torch.no_grad():
        for k, v in sd.items():
            dst = params[k]
            if v.is_floating_point() and v.dtype not in (torch.float8_e4m3fn, torch.float8_e5m2) \
                    and not k.endswith(".scale") and dst.dtype != v.dtype:
                v = v.to(dst.dtype)
            if dst.dtype in (torch.float8_e4m3fn, torch.float8_e5m2) and v.dtype == dst.dtype:
                dst.view(torch.uint8).copy_(v.contiguous().view(torch.uint8))
            else:
                dst.copy_(v)
    return sd
