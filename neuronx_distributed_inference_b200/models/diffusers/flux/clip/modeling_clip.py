"""CLIP text encoder (pooled prompt embedding of FLUX; reference models/diffusers/flux/clip/modeling_clip.py).  Checkpoint
layout: transformers ``CLIPTextModel``."""
from __future__ import annotations

import torch
import torch.nn as nn

from .....modules.vision import VisionAttention, VisionMLP


class CLIPLayer(nn.Module):
    def __init__(self, c, dtype, device):
        super().__init__()
        self.layer_norm1 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps, dtype=dtype, device=device)
        self.layer_norm2 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps, dtype=dtype, device=device)
        self.self_attn = VisionAttention(c.hidden_size, c.num_attention_heads, True, dtype, device)
        self.mlp = VisionMLP(c.hidden_size, c.intermediate_size, c.hidden_act, True, False, dtype, device)

    def forward(self, x, mask):
        x = x + self.self_attn(self.layer_norm1(x), mask=mask)
        return x + self.mlp(self.layer_norm2(x))


class NeuronClipTextModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        c, dt = config, config.neuron_config.torch_dtype
        self.config = config
        self.token_embedding = nn.Embedding(c.vocab_size, c.hidden_size, dtype=dt, device=device)
        self.position_embedding = nn.Embedding(c.max_position_embeddings, c.hidden_size, dtype=dt, device=device)
        self.layers = nn.ModuleList([CLIPLayer(c, dt, device) for _ in range(c.num_hidden_layers)])
        self.final_layer_norm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps, dtype=dt, device=device)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, input_ids):
        """-> (last_hidden_state [B,T,H], pooled [B,H] = state at the EOS token)."""
        B, T = input_ids.shape
        x = self.token_embedding(input_ids) + self.position_embedding.weight[:T]
        causal = torch.ones(T, T, dtype=torch.bool, device=input_ids.device).tril().view(1, 1, T, T)
        for layer in self.layers:
            x = layer(x, causal)
        x = self.final_layer_norm(x)
        eos = getattr(self.config, "eos_token_id", 2)
        idx = input_ids.argmax(-1) if eos == 2 else (input_ids == eos).int().argmax(-1)
        return x, x[torch.arange(B, device=x.device), idx]


def convert_clip_state_dict(sd: dict, config) -> dict:
    from ....state_dict_utils import fuse_qkv_and_gate_up
    out = {}
    for k, v in sd.items():
        k = k.replace("text_model.", "").replace("encoder.layers.", "layers.").replace("embeddings.", "")
        k = k.replace(".self_attn.out_proj.", ".self_attn.o_proj.")
        out[k] = v
    out = fuse_qkv_and_gate_up(out, config.num_hidden_layers, fuse_mlp=False)
    return {k: v for k, v in out.items() if "position_ids" not in k}
