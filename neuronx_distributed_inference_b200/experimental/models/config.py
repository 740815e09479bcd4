"""Architecture hyper-parameters of the experimental functional models (reference experimental/models/config.py:9-68): one flat
dataclass shared by Llama-3 and Llama-4, loadable from the ``params.json`` of a Meta checkpoint or from a Hugging Face
``config.json`` (text section)."""
from __future__ import annotations

import json
from dataclasses import dataclass, fields
from typing import List, Optional

import torch


@dataclass
class Config:
    dim: int = 4096
    n_layers: int = 32
    n_heads: int = 32
    n_kv_heads: int = 8
    head_dim: Optional[int] = None
    vocab_size: int = 128256
    ffn_dim: int = 14336                       # dense MLP width (Llama-4: ``intermediate_size_mlp``)
    norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = None
    max_batch_size: int = 2
    max_seq_len: int = 2048
    dtype: torch.dtype = torch.bfloat16
    # Llama-4
    n_experts: int = 0
    top_k: int = 1
    moe_ffn_dim: int = 8192                    # expert / shared-expert width (``intermediate_size``)
    moe_layers: Optional[List[int]] = None
    nope_layers: Optional[List[int]] = None    # layers WITHOUT rotary embedding (global attention)
    attention_chunk_size: Optional[int] = 8192
    use_qk_norm: bool = True
    attn_temperature_tuning: bool = True
    floor_scale: float = 8192.0
    attn_scale: float = 0.1

    def __post_init__(self):
        if self.head_dim is None:
            self.head_dim = self.dim // self.n_heads

    @classmethod
    def from_hf(cls, path_or_dict, **overrides) -> "Config":
        d = path_or_dict
        if isinstance(d, str):
            with open(d) as f:
                d = json.load(f)
        d = d.get("text_config", d)
        n = d["num_hidden_layers"]
        step_moe = d.get("interleave_moe_layer_step", 1)
        nope = d.get("no_rope_layers")
        kw = dict(dim=d["hidden_size"], n_layers=n, n_heads=d["num_attention_heads"], n_kv_heads=d.get("num_key_value_heads", d["num_attention_heads"]),
                  head_dim=d.get("head_dim"), vocab_size=d["vocab_size"], ffn_dim=d.get("intermediate_size_mlp", d.get("intermediate_size")),
                  norm_eps=d.get("rms_norm_eps", 1e-5), rope_theta=(d.get("rope_parameters") or {}).get("rope_theta", d.get("rope_theta", 500000.0)),
                  n_experts=d.get("num_local_experts", 0), top_k=d.get("num_experts_per_tok", 1), moe_ffn_dim=d.get("intermediate_size", 0),
                  moe_layers=d.get("moe_layers") or (list(range(step_moe - 1, n, step_moe)) if d.get("num_local_experts") else []),
                  nope_layers=[i for i, r in enumerate(nope) if not r] if nope else [i for i in range(n) if (i + 1) % 4 == 0],
                  attention_chunk_size=d.get("attention_chunk_size"), use_qk_norm=d.get("use_qk_norm", True),
                  attn_temperature_tuning=bool(d.get("attn_temperature_tuning", True)), floor_scale=d.get("floor_scale", 8192.0),
                  attn_scale=d.get("attn_scale", 0.1))
        kw.update(overrides)
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in kw.items() if k in known})
