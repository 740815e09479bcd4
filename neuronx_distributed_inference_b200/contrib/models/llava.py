"""LLaVA-1.5 (reference contrib/models/llava-v1.5-7b): CLIP ViT tower (class token, learned positions, pre-LN, quick-GELU),
hidden state of the penultimate layer without the class token, 2-layer GELU projector, Llama / Vicuna decoder."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...models.image_to_text_model_base import ImageToTextInferenceConfig, NeuronBaseForImageToText
from ...models.llama.modeling_llama import NeuronLlamaModel
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from ...modules.vision import ACT, PatchEmbed, VisionAttention, VisionMLP


class LlavaInferenceConfig(ImageToTextInferenceConfig):
    def get_required_attributes(self):
        return ["text_config", "vision_config"]


class ClipVisionLayer(nn.Module):
    def __init__(self, vc, dtype, device):
        super().__init__()
        eps = getattr(vc, "layer_norm_eps", 1e-5)
        self.layer_norm1 = nn.LayerNorm(vc.hidden_size, eps=eps, dtype=dtype, device=device)
        self.layer_norm2 = nn.LayerNorm(vc.hidden_size, eps=eps, dtype=dtype, device=device)
        self.self_attn = VisionAttention(vc.hidden_size, vc.num_attention_heads, True, dtype, device)
        self.mlp = VisionMLP(vc.hidden_size, vc.intermediate_size, getattr(vc, "hidden_act", "quick_gelu"), True, False, dtype, device)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class NeuronLlavaVisionModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        vc = config.vision_config
        dt = vc.neuron_config.torch_dtype
        self.vc = vc
        n = (vc.image_size // vc.patch_size) ** 2
        self.patch_embedding = PatchEmbed(vc.num_channels * vc.patch_size ** 2, vc.hidden_size, False, dt, device)
        self.class_embedding = nn.Parameter(torch.zeros(vc.hidden_size, dtype=dt, device=device), requires_grad=False)
        self.position_embedding = nn.Embedding(n + 1, vc.hidden_size, dtype=dt, device=device)
        self.pre_layrnorm = nn.LayerNorm(vc.hidden_size, eps=getattr(vc, "layer_norm_eps", 1e-5), dtype=dt, device=device)
        self.layers = nn.ModuleList([ClipVisionLayer(vc, dt, device) for _ in range(vc.num_hidden_layers)])
        tc = config.get_text_config()
        bias = getattr(config, "multimodal_projector_bias", True)
        self.proj1 = nn.Linear(vc.hidden_size, tc.hidden_size, bias=bias, dtype=dt, device=device)
        self.proj2 = nn.Linear(tc.hidden_size, tc.hidden_size, bias=bias, dtype=dt, device=device)
        self.proj_act = getattr(config, "projector_hidden_act", "gelu")
        self.feature_layer = getattr(config, "vision_feature_layer", -2)
        self.drop_cls = getattr(config, "vision_feature_select_strategy", "default") == "default"
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, pixel_values):
        n, C, H, W = pixel_values.shape
        P = self.vc.patch_size
        x = pixel_values.reshape(n, C, H // P, P, W // P, P).permute(0, 2, 4, 1, 3, 5).reshape(n, -1, C * P * P)
        x = self.patch_embedding(x)
        x = torch.cat([self.class_embedding.view(1, 1, -1).expand(n, 1, -1), x], 1) + self.position_embedding.weight[: x.shape[1] + 1]
        x = self.pre_layrnorm(x)
        hs = [x]
        for layer in self.layers:
            x = layer(x)
            hs.append(x)
        feat = hs[self.feature_layer]
        if self.drop_cls:
            feat = feat[:, 1:]
        out = self.proj2(ACT[self.proj_act](self.proj1(feat)))
        return out.reshape(-1, out.shape[-1])


class NeuronLlavaForCausalLM(NeuronBaseForImageToText):
    _model_cls = NeuronLlamaModel
    _vision_cls = NeuronLlavaVisionModel
    text_prefix = "language_model."
    vision_prefix = "vision_tower."

    @classmethod
    def get_config_cls(cls):
        return LlavaInferenceConfig

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
        sd = {("vision_tower.proj1." + k.split("linear_1.")[1] if k.startswith("multi_modal_projector.linear_1.") else
               "vision_tower.proj2." + k.split("linear_2.")[1] if k.startswith("multi_modal_projector.linear_2.") else k): v for k, v in sd.items()}
        return super()._split_state_dict(sd)

    @classmethod
    def get_state_dict(cls, path, config):
        from ...modules.checkpoint import load_state_dict
        sd = {cls._strip(k): v for k, v in load_state_dict(path).items()}
        text = {k[len(cls.text_prefix):] if k.startswith(cls.text_prefix) else k: v for k, v in sd.items()
                if not k.startswith(cls.vision_prefix) and not k.startswith("multi_modal_projector.")}
        text = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in text.items()}
        text = cls.convert_hf_to_neuron_state_dict(text, config.get_text_config())
        if getattr(config, "tie_word_embeddings", False) and "lm_head.weight" not in text:
            cls.update_state_dict_for_tied_weights(text)
        out = {cls.text_prefix + k: v for k, v in text.items()}
        vis = {}
        for k, v in sd.items():
            if not k.startswith(cls.vision_prefix):
                continue
            k = k[len(cls.vision_prefix):].replace("vision_model.", "")
            k = (k.replace("embeddings.class_embedding", "class_embedding").replace("embeddings.position_embedding.", "position_embedding.")
                 .replace("embeddings.patch_embedding.weight", "patch_embedding.proj.weight").replace("encoder.layers.", "layers.")
                 .replace(".self_attn.out_proj.", ".self_attn.o_proj."))
            if k == "patch_embedding.proj.weight":
                v = v.reshape(v.shape[0], -1)
            if k.startswith("post_layernorm.") or "position_ids" in k:
                continue
            vis[k] = v
        vis = fuse_qkv_and_gate_up(vis, config.vision_config.num_hidden_layers, fuse_mlp=False)
        out.update({cls.vision_prefix + k: v for k, v in vis.items()})
        out.update({k: v for k, v in sd.items() if k.startswith("multi_modal_projector.")})
        return out

    def encode_images(self, pixel_values, **kw):
        return self.vision_encoder_model(pixel_values)
