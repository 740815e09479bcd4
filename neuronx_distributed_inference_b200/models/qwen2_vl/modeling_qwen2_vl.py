"""Qwen2-VL / Qwen2.5-VL-style image-to-text model: ViT with 2-D rotary + patch merger, Qwen2 text decoder with M-RoPE.

reference: models/qwen2_vl/modeling_qwen2_vl.py, modeling_qwen2_vl_text.py, modeling_qwen2_vl_vision.py (≈1350 LoC) on top of
``NeuronBaseForImageToText``.  Inputs follow the Hugging Face processor: ``pixel_values`` ``[n_patches, C*t*p*p]`` (flattened
patches in merge-window order), ``image_grid_thw`` ``[n_images, 3]``."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ...modules.rope import MRotaryEmbedding
from ...modules.vision import PatchEmbed, VisionAttention, VisionMLP
from ..image_to_text_model_base import ImageToTextInferenceConfig, NeuronBaseForImageToText
from ..llama.modeling_llama import rope_scaling_of, rope_theta_of
from ..qwen2.modeling_qwen2 import NeuronQwen2Model
from ..state_dict_utils import fuse_qkv_and_gate_up


class Qwen2VLInferenceConfig(ImageToTextInferenceConfig):
    def get_required_attributes(self):
        return ["text_config", "vision_config"]


def mrope_section_of(cfg):
    rs = rope_scaling_of(cfg) or getattr(cfg, "rope_parameters", None) or {}
    sec = rs.get("mrope_section") if isinstance(rs, dict) else None
    if sec is None:
        raise ValueError("M-RoPE needs rope_scaling/rope_parameters['mrope_section']")
    return list(sec)


class NeuronQwen2VLTextModel(NeuronQwen2Model):
    def make_rotary(self, config, device):
        return MRotaryEmbedding(config.head_dim, max(config.max_position_embeddings, config.neuron_config.seq_len),
                                rope_theta_of(config), mrope_section_of(config), device=device)


class Qwen2VLVisionBlock(nn.Module):
    def __init__(self, vc, dtype, device):
        super().__init__()
        self.norm1 = nn.LayerNorm(vc.embed_dim, eps=1e-6, dtype=dtype, device=device)
        self.norm2 = nn.LayerNorm(vc.embed_dim, eps=1e-6, dtype=dtype, device=device)
        self.attn = VisionAttention(vc.embed_dim, vc.num_heads, True, dtype, device)
        self.mlp = VisionMLP(vc.embed_dim, int(vc.embed_dim * vc.mlp_ratio), getattr(vc, "hidden_act", "quick_gelu"), True, False,
                             dtype, device)

    def forward(self, x, cos, sin, seg):
        x = x + self.attn(self.norm1(x), cos, sin, seg)
        return x + self.mlp(self.norm2(x))


class PatchMerger(nn.Module):
    def __init__(self, out_dim, ctx_dim, merge, dtype, device):
        super().__init__()
        self.hidden = ctx_dim * merge * merge
        self.ln_q = nn.LayerNorm(ctx_dim, eps=1e-6, dtype=dtype, device=device)
        self.mlp = nn.Sequential(nn.Linear(self.hidden, self.hidden, dtype=dtype, device=device), nn.GELU(),
                                 nn.Linear(self.hidden, out_dim, dtype=dtype, device=device))

    def forward(self, x):
        return self.mlp(self.ln_q(x).view(-1, self.hidden))


class NeuronQwen2VLVisionModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        vc = config.vision_config
        dt = vc.neuron_config.torch_dtype
        self.vc = vc
        self.merge = vc.spatial_merge_size
        self.patch_embed = PatchEmbed(vc.in_channels * vc.temporal_patch_size * vc.patch_size ** 2, vc.embed_dim, False, dt, device)
        self.blocks = nn.ModuleList([Qwen2VLVisionBlock(vc, dt, device) for _ in range(vc.depth)])
        self.merger = PatchMerger(vc.hidden_size, vc.embed_dim, self.merge, dt, device)
        self.head_dim = vc.embed_dim // vc.num_heads
        for p in self.parameters():
            p.requires_grad_(False)

    def rot_pos(self, grid_thw: torch.Tensor, device):
        m = self.merge
        ids, seg, s = [], [], 0
        for t, h, w in grid_thw.tolist():
            hp = torch.arange(h).view(h, 1).expand(h, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
            wp = torch.arange(w).view(1, w).expand(h, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
            ids.append(torch.stack([hp, wp], -1).repeat(t, 1))
            for f in range(t):
                seg.append(torch.full((h * w,), s, dtype=torch.int32))
                s += 1
        ids = torch.cat(ids).to(device)
        dim = self.head_dim // 2
        inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device) / dim))
        fr = torch.outer(torch.arange(int(grid_thw[:, 1:].max()), dtype=torch.float32, device=device), inv)
        emb = fr[ids].flatten(1)                                   # [N, head_dim/2]
        emb = torch.cat([emb, emb], -1)
        return emb.cos(), emb.sin(), torch.cat(seg).to(device)

    def forward(self, pixel_values: torch.Tensor, image_grid_thw: torch.Tensor):
        x = self.patch_embed(pixel_values)                         # [N, C]
        cos, sin, seg = self.rot_pos(image_grid_thw.cpu(), x.device)
        x, cos, sin, seg = x.unsqueeze(0), cos.unsqueeze(0), sin.unsqueeze(0), seg.unsqueeze(0)
        for blk in self.blocks:
            x = blk(x, cos, sin, seg)
        return self.merger(x.squeeze(0))                           # [N / merge^2, H_text]


def get_rope_index(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], image_grid_thw: Optional[torch.Tensor],
                   image_token_id: int, merge: int, video_grid_thw=None, video_token_id=None) -> torch.Tensor:
    """3-D (t, h, w) positions: text tokens advance all three axes together, the tokens of an image take the coordinates of
    their merged patch offset by the running position (HF ``get_rope_index``).  -> [3,B,T] long."""
    B, T = input_ids.shape
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    out = torch.ones(3, B, T, dtype=torch.long)
    grids = [] if image_grid_thw is None else [(g, image_token_id) for g in image_grid_thw.tolist()]
    vgrids = [] if video_grid_thw is None else [(g, video_token_id) for g in video_grid_thw.tolist()]
    gi = vi = 0
    for b in range(B):
        valid = attention_mask[b].bool()
        toks = input_ids[b][valid].tolist()
        pos, st, cur = [], 0, 0
        while st < len(toks):
            nxt_img = toks.index(image_token_id, st) if image_token_id in toks[st:] and gi < len(grids) else None
            nxt_vid = toks.index(video_token_id, st) if (video_token_id is not None and video_token_id in toks[st:]
                                                         and vi < len(vgrids)) else None
            cands = [c for c in (nxt_img, nxt_vid) if c is not None]
            if not cands:
                n = len(toks) - st
                pos.append(torch.arange(n).view(1, n).expand(3, n) + cur)
                st = len(toks)
                break
            ed = min(cands)
            if ed == nxt_img:
                (t, h, w), _ = grids[gi]
                gi += 1
            else:
                (t, h, w), _ = vgrids[vi]
                vi += 1
            n = ed - st
            if n:
                pos.append(torch.arange(n).view(1, n).expand(3, n) + cur)
                cur += n
            gh, gw = h // merge, w // merge
            tt = torch.arange(t).view(t, 1).expand(t, gh * gw).flatten()
            hh = torch.arange(gh).view(1, gh, 1).expand(t, gh, gw).flatten()
            ww = torch.arange(gw).view(1, 1, gw).expand(t, gh, gw).flatten()
            pos.append(torch.stack([tt, hh, ww]) + cur)
            cur = int(pos[-1].max()) + 1
            st = ed + t * gh * gw
        p = torch.cat(pos, 1) if pos else torch.zeros(3, 0, dtype=torch.long)
        out[:, b, valid] = p[:, : int(valid.sum())]
    return out


class NeuronQwen2VLForCausalLM(NeuronBaseForImageToText):
    _model_cls = NeuronQwen2VLTextModel
    _vision_cls = NeuronQwen2VLVisionModel
    text_prefix = "language_model."
    vision_prefix = "visual."

    @classmethod
    def get_config_cls(cls):
        return Qwen2VLInferenceConfig

    @staticmethod
    def load_hf_model(model_path):
        from transformers import AutoModelForImageTextToText
        return AutoModelForImageTextToText.from_pretrained(model_path)

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        return fuse_qkv_and_gate_up(sd, config.num_hidden_layers)

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        sd["lm_head.weight"] = sd["embed_tokens.weight"].clone()

    @staticmethod
    def convert_hf_to_neuron_vision_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k == "patch_embed.proj.weight":
                v = v.reshape(v.shape[0], -1)
            k = k.replace(".attn.qkv.", ".attn.qkv_proj.").replace(".attn.proj.", ".attn.o_proj.")
            out[k] = v
        return out

    def get_rotary_position_ids(self, input_ids, attention_mask, image_grid_thw=None, video_grid_thw=None, **kw):
        if image_grid_thw is None and video_grid_thw is None:
            return None
        return get_rope_index(input_ids.cpu(), None if attention_mask is None else attention_mask.cpu(), image_grid_thw,
                              self.config.image_token_id, self.config.vision_config.spatial_merge_size, video_grid_thw,
                              getattr(self.config, "video_token_id", None))

    def encode_images(self, pixel_values, image_grid_thw=None, **kw):
        return self.vision_encoder_model(pixel_values, image_grid_thw)
