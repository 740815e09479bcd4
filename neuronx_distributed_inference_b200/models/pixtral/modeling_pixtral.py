"""Pixtral (Llava-style): Pixtral ViT (2-D RoPE, RMSNorm, SwiGLU, variable image sizes) + MLP projector + Mistral decoder.

reference: models/pixtral/modeling_pixtral.py + modeling_pixtral_vision.py (≈1109 LoC) on ``NeuronBaseForImageToText``.
Inputs follow the HF processor: ``pixel_values`` ``[n_images, C, H_max, W_max]`` and ``image_sizes`` ``[n_images, 2]``; all the
patches of all images run as one packed sequence with block-diagonal attention."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...modules.norm import RMSNorm
from ...modules.vision import ACT, PatchEmbed, VisionAttention, VisionMLP
from ..image_to_text_model_base import ImageToTextInferenceConfig, NeuronBaseForImageToText
from ..mistral.modeling_mistral import NeuronMistralModel
from ..state_dict_utils import fuse_qkv_and_gate_up


class PixtralInferenceConfig(ImageToTextInferenceConfig):
    def get_required_attributes(self):
        return ["text_config", "vision_config"]


class PixtralVisionLayer(nn.Module):
    def __init__(self, vc, dtype, device):
        super().__init__()
        hd = getattr(vc, "head_dim", None) or vc.hidden_size // vc.num_attention_heads
        self.attention_norm = RMSNorm(vc.hidden_size, 1e-5, dtype, device=device)
        self.ffn_norm = RMSNorm(vc.hidden_size, 1e-5, dtype, device=device)
        self.attention = VisionAttention(vc.hidden_size, vc.num_attention_heads, False, dtype, device, head_dim=hd)
        self.feed_forward = VisionMLP(vc.hidden_size, vc.intermediate_size, getattr(vc, "hidden_act", "gelu"), False, True, dtype, device)

    def forward(self, x, cos, sin, seg):
        x = x + self.attention(self.attention_norm(x), cos, sin, seg)
        return x + self.feed_forward(self.ffn_norm(x))


class NeuronPixtralVisionModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        vc = config.vision_config
        dt = vc.neuron_config.torch_dtype
        self.vc = vc
        self.patch = vc.patch_size
        self.patch_conv = PatchEmbed(vc.num_channels * vc.patch_size ** 2, vc.hidden_size, False, dt, device)
        self.ln_pre = RMSNorm(vc.hidden_size, 1e-5, dt, device=device)
        self.layers = nn.ModuleList([PixtralVisionLayer(vc, dt, device) for _ in range(vc.num_hidden_layers)])
        tc = config.get_text_config()
        self.proj1 = nn.Linear(vc.hidden_size, tc.hidden_size, bias=getattr(config, "multimodal_projector_bias", True), dtype=dt, device=device)
        self.proj2 = nn.Linear(tc.hidden_size, tc.hidden_size, bias=getattr(config, "multimodal_projector_bias", True), dtype=dt, device=device)
        self.proj_act = getattr(config, "projector_hidden_act", "gelu")
        self.feature_layer = getattr(config, "vision_feature_layer", -1)
        hd = getattr(vc, "head_dim", None) or vc.hidden_size // vc.num_attention_heads
        rp = getattr(vc, "rope_parameters", None) or {}
        base = float(rp.get("rope_theta", getattr(vc, "rope_theta", 10000.0)))
        side = vc.image_size // vc.patch_size
        fr = 1.0 / (base ** (torch.arange(0, hd, 2).float() / hd))
        fh = torch.outer(torch.arange(side).float(), fr[::2])
        fw = torch.outer(torch.arange(side).float(), fr[1::2])
        table = torch.cat([fh[:, None, :].repeat(1, side, 1), fw[None, :, :].repeat(side, 1, 1)], -1).reshape(-1, hd // 2)
        self.register_buffer("rope_table", torch.cat([table, table], -1).to(device), persistent=False)
        self.side = side
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, pixel_values: torch.Tensor, image_sizes=None):
        feat, _ = self.features(pixel_values, image_sizes)
        return self.proj2(ACT[self.proj_act](self.proj1(feat)))                 # [n_patches_total, H_text]

    def features(self, pixel_values: torch.Tensor, image_sizes=None):
        """-> (selected tower features [n_patches_total, H_vision(*n_feature_layers)], per-image patch grids [(gh, gw)])."""
        n, C, H, W = pixel_values.shape
        P = self.patch
        if image_sizes is None:
            image_sizes = [(H, W)] * n
        sizes = [(int(h) // P, int(w) // P) for h, w in (image_sizes.tolist() if torch.is_tensor(image_sizes) else image_sizes)]
        seqs, pos, seg = [], [], []
        for i, (gh, gw) in enumerate(sizes):
            img = pixel_values[i, :, : gh * P, : gw * P]
            patches = img.reshape(C, gh, P, gw, P).permute(1, 3, 0, 2, 4).reshape(gh * gw, C * P * P)
            seqs.append(patches)
            hh = torch.arange(gh).view(gh, 1).expand(gh, gw).reshape(-1)
            ww = torch.arange(gw).view(1, gw).expand(gh, gw).reshape(-1)
            pos.append(hh * self.side + ww)
            seg.append(torch.full((gh * gw,), i, dtype=torch.int32))
        dev = pixel_values.device
        x = self.ln_pre(self.patch_conv(torch.cat(seqs))).unsqueeze(0)
        emb = self.rope_table[torch.cat(pos).to(dev)]
        cos, sin, seg = emb.cos().unsqueeze(0), emb.sin().unsqueeze(0), torch.cat(seg).to(dev).unsqueeze(0)
        hs = [x]
        for layer in self.layers:
            x = layer(x, cos, sin, seg)
            hs.append(x)
        feat = hs[self.feature_layer] if isinstance(self.feature_layer, int) else torch.cat([hs[i] for i in self.feature_layer], -1)
        return feat.squeeze(0), sizes


class NeuronPixtralForCausalLM(NeuronBaseForImageToText):
    _model_cls = NeuronMistralModel
    _vision_cls = NeuronPixtralVisionModel
    text_prefix = "language_model."
    vision_prefix = "vision_tower."

    @classmethod
    def get_config_cls(cls):
        return PixtralInferenceConfig

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

    def _split_state_dict(self, sd):
        # the projector lives beside the tower in the checkpoint; it belongs to the vision module here
        sd = {("vision_tower.proj1." + k.split("linear_1.")[1] if k.startswith("multi_modal_projector.linear_1.") else
               "vision_tower.proj2." + k.split("linear_2.")[1] if k.startswith("multi_modal_projector.linear_2.") else k): v
              for k, v in sd.items()}
        return super()._split_state_dict(sd)

    @classmethod
    def get_state_dict(cls, path, config):
        from ...modules.checkpoint import load_state_dict
        sd = {cls._strip(k): v for k, v in load_state_dict(path).items()}
        text = {k[len(cls.text_prefix):] if k.startswith(cls.text_prefix) else k: v for k, v in sd.items()
                if not k.startswith(cls.vision_prefix) and not k.startswith("multi_modal_projector.")}
        text = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in text.items()}   # legacy llava layout
        text = cls.convert_hf_to_neuron_state_dict(text, config.get_text_config())
        if getattr(config, "tie_word_embeddings", False) and "lm_head.weight" not in text:
            cls.update_state_dict_for_tied_weights(text)
        out = {cls.text_prefix + k: v for k, v in text.items()}
        vis = {k[len(cls.vision_prefix):]: v for k, v in sd.items() if k.startswith(cls.vision_prefix)}
        n_layers = config.vision_config.num_hidden_layers
        vis = {k.replace("transformer.layers.", "layers."): v for k, v in vis.items()}
        vis = fuse_qkv_and_gate_up(vis, n_layers, attn="attention", mlp="feed_forward")
        vis = {k.replace(".feed_forward.down_proj.", ".feed_forward.fc2.").replace("patch_conv.weight", "patch_conv.proj.weight"): v
               for k, v in vis.items()}
        vis["patch_conv.proj.weight"] = vis["patch_conv.proj.weight"].reshape(vis["patch_conv.proj.weight"].shape[0], -1)
        out.update({cls.vision_prefix + k: v for k, v in vis.items()})
        out.update({k: v for k, v in sd.items() if k.startswith("multi_modal_projector.")})
        return out

    def encode_images(self, pixel_values, image_sizes=None, **kw):
        return self.vision_encoder_model(pixel_values, image_sizes)
