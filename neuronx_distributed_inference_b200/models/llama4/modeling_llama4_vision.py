"""Llama-4 vision tower (reference models/llama4/modeling_llama4_vision.py:1-1214): unfold-convolution patch embedding, class
token appended *last*, learned position table, ViT layers with interleaved 2-D rotary (x half / y half of each head), pixel-
shuffle adapter MLP and the linear multimodal projector."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...modules.vision import PatchEmbed, VisionAttention, VisionMLP


def pixel_shuffle(x: torch.Tensor, ratio: float) -> torch.Tensor:
    B, N, C = x.shape
    s = int(math.sqrt(N))
    x = x.view(B, s, int(s * ratio), int(C / ratio)).permute(0, 2, 1, 3).contiguous()
    x = x.view(B, int(s * ratio), int(s * ratio), int(C / ratio ** 2)).permute(0, 2, 1, 3).contiguous()
    return x.view(B, -1, x.shape[-1])


class Llama4VisionLayer(nn.Module):
    def __init__(self, vc, dtype, device):
        super().__init__()
        self.input_layernorm = nn.LayerNorm(vc.hidden_size, dtype=dtype, device=device)
        self.post_attention_layernorm = nn.LayerNorm(vc.hidden_size, dtype=dtype, device=device)
        self.self_attn = VisionAttention(vc.hidden_size, vc.num_attention_heads, True, dtype, device)
        self.mlp = VisionMLP(vc.hidden_size, vc.intermediate_size, "gelu", True, False, dtype, device)

    def forward(self, x, cos, sin):
        x = x + self.self_attn(self.input_layernorm(x), cos, sin)
        return x + self.mlp(self.post_attention_layernorm(x))


class NeuronLlama4VisionModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        vc = config.vision_config
        dt = vc.neuron_config.torch_dtype
        self.vc = vc
        H = vc.hidden_size
        side = vc.image_size // vc.patch_size
        self.num_patches = side * side + 1
        self.patch_embedding = PatchEmbed(vc.num_channels * vc.patch_size ** 2, H, False, dt, device)
        self.class_embedding = nn.Parameter(torch.zeros(H, dtype=dt, device=device), requires_grad=False)
        self.positional_embedding_vlm = nn.Parameter(torch.zeros(self.num_patches, H, dtype=dt, device=device), requires_grad=False)
        self.layernorm_pre = nn.LayerNorm(H, dtype=dt, device=device)
        self.layernorm_post = nn.LayerNorm(H, dtype=dt, device=device)
        self.layers = nn.ModuleList([Llama4VisionLayer(vc, dt, device) for _ in range(vc.num_hidden_layers)])
        self.adapter_fc1 = nn.Linear(vc.intermediate_size, vc.projector_input_dim, bias=False, dtype=dt, device=device)
        self.adapter_fc2 = nn.Linear(vc.projector_output_dim, vc.projector_output_dim, bias=False, dtype=dt, device=device)
        self.projector = nn.Linear(vc.vision_output_dim, config.get_text_config().hidden_size, bias=False, dtype=dt, device=device)
        # rotary table: per token (x+1, y+1) angles, class token unrotated
        rp = getattr(vc, "rope_parameters", None) or {}
        theta = float(rp.get("rope_theta", getattr(vc, "rope_theta", 10000.0)))
        idx = torch.arange(side * side, dtype=torch.int32).view(-1, 1)
        idx = torch.cat([idx, idx[:1]], 0)
        idx[-1, -1] = -2
        fx, fy = idx % side, idx // side
        fd = H // vc.num_attention_heads // 2
        rf = 1.0 / (theta ** (torch.arange(0, fd, 2)[: fd // 2].float() / fd))
        ax = ((fx + 1)[..., None] * rf[None, None, :]).repeat_interleave(2, -1)
        ay = ((fy + 1)[..., None] * rf[None, None, :]).repeat_interleave(2, -1)
        fr = torch.cat([ax, ay], -1).float()[..., ::2].masked_fill(idx.reshape(-1, 1, 1) < 0, 0).squeeze(1)   # [N, hd/2]
        self.register_buffer("rope_angles", fr.to(device), persistent=False)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, pixel_values):
        n, C, Hh, Ww = pixel_values.shape
        P = self.vc.patch_size
        # nn.Unfold ordering: feature index = c*P*P + ph*P + pw, patches row-major
        x = pixel_values.reshape(n, C, Hh // P, P, Ww // P, P).permute(0, 2, 4, 1, 3, 5).reshape(n, -1, C * P * P)
        x = self.patch_embedding(x)
        x = torch.cat([x, self.class_embedding.view(1, 1, -1).expand(n, 1, -1)], 1)
        x = self.layernorm_pre(x + self.positional_embedding_vlm.to(x.dtype))
        cos = self.rope_angles.cos().unsqueeze(0).expand(n, -1, -1)
        sin = self.rope_angles.sin().unsqueeze(0).expand(n, -1, -1)
        for layer in self.layers:
            x = layer(x, cos, sin)
        x = self.layernorm_post(x)[:, :-1]
        x = pixel_shuffle(x, self.vc.pixel_shuffle_ratio)
        g = nn.functional.gelu
        x = g(self.adapter_fc2(g(self.adapter_fc1(x))))
        return self.projector(x.reshape(-1, x.shape[-1]))                       # [n_tiles * tokens, H_text]
