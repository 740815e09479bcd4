"""Sharded Hugging Face safetensors -> one state dict (reference experimental/core/checkpoint.py:11-63): reads
``model.safetensors.index.json`` when present, otherwise every ``*.safetensors`` file of the directory."""
from __future__ import annotations

import glob
import json
import os
from typing import Dict

import torch


def load_hf_safetensors_sharded(state_dict_dir: str) -> Dict[str, torch.Tensor]:
    from safetensors.torch import load_file
    index = os.path.join(state_dict_dir, "model.safetensors.index.json")
    if os.path.exists(index):
        with open(index) as f:
            files = sorted(set(json.load(f)["weight_map"].values()))
    else:
        files = sorted(os.path.basename(p) for p in glob.glob(os.path.join(state_dict_dir, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no safetensors files under {state_dict_dir}")
    sd: Dict[str, torch.Tensor] = {}
    for name in files:
        part = load_file(os.path.join(state_dict_dir, name))
        dup = set(part) & set(sd)
        if dup:
            raise ValueError(f"tensors present in more than one shard: {sorted(dup)[:4]}")
        sd.update(part)
    return sd
