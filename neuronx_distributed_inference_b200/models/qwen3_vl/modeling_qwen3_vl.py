"""Qwen3-VL: ViT with interpolated learned position embeddings + 2-D rotary, *deepstack* mergers (features of intermediate
vision layers are added to the text residual stream after the first decoder layers), Qwen3 text decoder (per-head q/k RMSNorm)
with interleaved M-RoPE.

reference: models/qwen3_vl/modeling_qwen3_vl.py, modeling_qwen3_vl_text.py, modeling_qwen3_vl_vision.py (≈2318 LoC)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...modules.rope import MRotaryEmbedding
from ...modules.vision import PatchEmbed, VisionAttention, VisionMLP
from ..image_to_text_model_base import NeuronBaseForImageToText
from ..llama.modeling_llama import rope_theta_of
from ..qwen2_vl.modeling_qwen2_vl import Qwen2VLInferenceConfig, get_rope_index, mrope_section_of
from ..qwen3.modeling_qwen3 import NeuronQwen3Model
from ..state_dict_utils import fuse_qkv_and_gate_up


class Qwen3VLInferenceConfig(Qwen2VLInferenceConfig):
    pass


class NeuronQwen3VLTextModel(NeuronQwen3Model):
    def make_rotary(self, config, device):
        rp = getattr(config, "rope_parameters", None) or getattr(config, "rope_scaling", None) or {}
        return MRotaryEmbedding(config.head_dim, max(config.max_position_embeddings, config.neuron_config.seq_len),
                                rope_theta_of(config), mrope_section_of(config), device=device,
                                interleaved=bool(rp.get("mrope_interleaved", True)))


class Qwen3VLVisionBlock(nn.Module):
    def __init__(self, vc, dtype, device):
        super().__init__()
        self.norm1 = nn.LayerNorm(vc.hidden_size, eps=1e-6, dtype=dtype, device=device)
        self.norm2 = nn.LayerNorm(vc.hidden_size, eps=1e-6, dtype=dtype, device=device)
        self.attn = VisionAttention(vc.hidden_size, vc.num_heads, True, dtype, device)
        self.mlp = VisionMLP(vc.hidden_size, vc.intermediate_size, getattr(vc, "hidden_act", "gelu_pytorch_tanh"), True, False, dtype, device)

    def forward(self, x, cos, sin, seg):
        x = x + self.attn(self.norm1(x), cos, sin, seg)
        return x + self.mlp(self.norm2(x))


class Qwen3VLPatchMerger(nn.Module):
    def __init__(self, vc, postshuffle: bool, dtype, device):
        super().__init__()
        self.hidden = vc.hidden_size * vc.spatial_merge_size ** 2
        self.postshuffle = postshuffle
        self.norm = nn.LayerNorm(self.hidden if postshuffle else vc.hidden_size, eps=1e-6, dtype=dtype, device=device)
        self.linear_fc1 = nn.Linear(self.hidden, self.hidden, dtype=dtype, device=device)
        self.linear_fc2 = nn.Linear(self.hidden, vc.out_hidden_size, dtype=dtype, device=device)

    def forward(self, x):
        x = self.norm(x.view(-1, self.hidden) if self.postshuffle else x).view(-1, self.hidden)
        return self.linear_fc2(nn.functional.gelu(self.linear_fc1(x)))


class NeuronQwen3VLVisionModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        vc = config.vision_config
        dt = vc.neuron_config.torch_dtype
        self.vc, self.merge = vc, vc.spatial_merge_size
        self.patch_embed = PatchEmbed(vc.in_channels * vc.temporal_patch_size * vc.patch_size ** 2, vc.hidden_size, True, dt, device)
        self.pos_embed = nn.Embedding(vc.num_position_embeddings, vc.hidden_size, dtype=dt, device=device)
        self.side = int(vc.num_position_embeddings ** 0.5)
        self.blocks = nn.ModuleList([Qwen3VLVisionBlock(vc, dt, device) for _ in range(vc.depth)])
        self.merger = Qwen3VLPatchMerger(vc, False, dt, device)
        self.deepstack_idx = list(getattr(vc, "deepstack_visual_indexes", []))
        self.deepstack_merger_list = nn.ModuleList([Qwen3VLPatchMerger(vc, True, dt, device) for _ in self.deepstack_idx])
        self.head_dim = vc.hidden_size // vc.num_heads
        for p in self.parameters():
            p.requires_grad_(False)

    def interp_pos(self, grid_thw, device):
        """Bilinear interpolation of the learned ``side x side`` position table to each image grid, emitted in merge-window
        order (HF ``fast_pos_embed_interpolate``)."""
        out, m, S = [], self.merge, self.side
        W = self.pos_embed.weight
        for t, h, w in grid_thw.tolist():
            hi, wi = torch.linspace(0, S - 1, h), torch.linspace(0, S - 1, w)
            hf, wf = hi.int(), wi.int()
            hc, wc = (hf + 1).clip(max=S - 1), (wf + 1).clip(max=S - 1)
            dh, dw = (hi - hf).view(h, 1), (wi - wf).view(1, w)
            idx = [(hf.view(h, 1) * S + wf.view(1, w)), (hf.view(h, 1) * S + wc.view(1, w)),
                   (hc.view(h, 1) * S + wf.view(1, w)), (hc.view(h, 1) * S + wc.view(1, w))]
            wts = [(1 - dh) * (1 - dw), (1 - dh) * dw, dh * (1 - dw), dh * dw]
            pe = sum(W[i.long().to(device).flatten()] * wt.flatten().to(device, W.dtype).unsqueeze(-1) for i, wt in zip(idx, wts))
            pe = pe.repeat(t, 1).view(t, h // m, m, w // m, m, -1).permute(0, 1, 3, 2, 4, 5).flatten(0, 4)
            out.append(pe)
        return torch.cat(out)

    def rot_pos(self, grid_thw, device):
        m, ids, seg, s = self.merge, [], [], 0
        for t, h, w in grid_thw.tolist():
            hp = torch.arange(h).view(h, 1).expand(h, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
            wp = torch.arange(w).view(1, w).expand(h, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
            ids.append(torch.stack([hp, wp], -1).repeat(t, 1))
            for _ in range(t):
                seg.append(torch.full((h * w,), s, dtype=torch.int32))
                s += 1
        ids = torch.cat(ids).to(device)
        dim = self.head_dim // 2
        inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device) / dim))
        fr = torch.outer(torch.arange(int(grid_thw[:, 1:].max()), dtype=torch.float32, device=device), inv)
        emb = fr[ids].flatten(1)
        emb = torch.cat([emb, emb], -1)
        return emb.cos(), emb.sin(), torch.cat(seg).to(device)

    def forward(self, pixel_values, image_grid_thw):
        g = image_grid_thw.cpu()
        x = self.patch_embed(pixel_values)
        x = x + self.interp_pos(g, x.device).to(x.dtype)
        cos, sin, seg = self.rot_pos(g, x.device)
        x, cos, sin, seg = x.unsqueeze(0), cos.unsqueeze(0), sin.unsqueeze(0), seg.unsqueeze(0)
        deep = []
        for i, blk in enumerate(self.blocks):
            x = blk(x, cos, sin, seg)
            if i in self.deepstack_idx:
                deep.append(self.deepstack_merger_list[self.deepstack_idx.index(i)](x.squeeze(0)))
        return self.merger(x.squeeze(0)), deep


class NeuronQwen3VLForCausalLM(NeuronBaseForImageToText):
    _model_cls = NeuronQwen3VLTextModel
    _vision_cls = NeuronQwen3VLVisionModel
    text_prefix = "language_model."
    vision_prefix = "visual."

    @classmethod
    def get_config_cls(cls):
        return Qwen3VLInferenceConfig

    @staticmethod
    def load_hf_model(model_path):
        from transformers import AutoModelForImageTextToText
        return AutoModelForImageTextToText.from_pretrained(model_path)

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        from ..qwen3.modeling_qwen3 import NeuronQwen3ForCausalLM
        return NeuronQwen3ForCausalLM.convert_hf_to_neuron_state_dict(sd, config)

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        sd["lm_head.weight"] = sd["embed_tokens.weight"].clone()

    @staticmethod
    def convert_hf_to_neuron_vision_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k == "patch_embed.proj.weight":
                v = v.reshape(v.shape[0], -1)
            k = (k.replace(".attn.qkv.", ".attn.qkv_proj.").replace(".attn.proj.", ".attn.o_proj.")
                 .replace(".mlp.linear_fc1.", ".mlp.fc1.").replace(".mlp.linear_fc2.", ".mlp.fc2."))
            out[k] = v
        return out

    def get_rotary_position_ids(self, input_ids, attention_mask, image_grid_thw=None, video_grid_thw=None, **kw):
        if image_grid_thw is None and video_grid_thw is None:
            return None
        return get_rope_index(input_ids.cpu(), None if attention_mask is None else attention_mask.cpu(), image_grid_thw,
                              self.config.image_token_id, self.config.vision_config.spatial_merge_size, video_grid_thw,
                              getattr(self.config, "video_token_id", None))

    def encode_images(self, pixel_values, image_grid_thw=None, **kw):
        emb, deep = self.vision_encoder_model(pixel_values, image_grid_thw)
        self._deepstack = deep
        return emb

    def forward(self, input_ids, *a, **kw):
        self._deepstack = None
        pix = kw.get("pixel_values")
        if pix is not None and kw.get("vision_embeddings") is None and input_ids.shape[-1] > 1:
            vis_kw = {k: kw[k] for k in self.vision_kwargs if k in kw}
            kw["vision_embeddings"] = self.encode_images(pix, **vis_kw)
            kw["deepstack_embeds"] = self._deepstack
        return super().forward(input_ids, *a, **kw)
