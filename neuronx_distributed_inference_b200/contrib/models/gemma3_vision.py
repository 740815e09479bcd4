"""Gemma-3 vision-language models (4B / 12B / 27B): SigLIP ViT tower (no class token, learned positions, pre-LN, GELU-tanh MLP, final
LayerNorm), the Gemma-3 projector (average-pool the patch grid down to ``mm_tokens_per_image`` soft tokens, offset RMSNorm, one
matmul) and the Gemma-3 text decoder, in which the soft tokens of one image attend to each other bidirectionally during prefill.
reference port: contrib/models/gemma3-vision/src/gemma3_vision."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...models.gemma3.modeling_gemma3 import NeuronGemma3ForCausalLM, NeuronGemma3Model
from ...models.image_to_text_model_base import ImageToTextInferenceConfig, NeuronBaseForImageToText
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from ...modules.norm import RMSNorm
from ...modules.vision import PatchEmbed
from .llava import ClipVisionLayer


class Gemma3VisionInferenceConfig(ImageToTextInferenceConfig):
    def get_required_attributes(self):
        return ["text_config", "vision_config"]

    def add_derived_config(self):
        super().add_derived_config()
        tc = self.text_config
        act = getattr(tc, "hidden_activation", None) or "gelu_pytorch_tanh"
        object.__setattr__(tc, "hidden_act", act)
        if getattr(tc, "num_key_value_heads", None) is None:
            object.__setattr__(tc, "num_key_value_heads", tc.num_attention_heads)
        if not hasattr(self.vision_config, "hidden_act"):
            object.__setattr__(self.vision_config, "hidden_act", "gelu_pytorch_tanh")


class NeuronGemma3VisionModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        vc, tc = config.vision_config, config.get_text_config()
        dt = vc.neuron_config.torch_dtype
        self.vc = vc
        self.side = vc.image_size // vc.patch_size
        eps = getattr(vc, "layer_norm_eps", 1e-6)
        self.patch_embedding = PatchEmbed(vc.num_channels * vc.patch_size ** 2, vc.hidden_size, True, dt, device)
        self.position_embedding = nn.Embedding(self.side ** 2, vc.hidden_size, dtype=dt, device=device)
        self.layers = nn.ModuleList([ClipVisionLayer(vc, dt, device) for _ in range(vc.num_hidden_layers)])
        self.post_layernorm = nn.LayerNorm(vc.hidden_size, eps=eps, dtype=dt, device=device)
        self.mm_soft_emb_norm = RMSNorm(vc.hidden_size, eps, dt, offset=1.0, device=device)
        self.mm_input_projection_weight = nn.Parameter(torch.zeros(vc.hidden_size, tc.hidden_size, dtype=dt, device=device), requires_grad=False)
        self.pool = self.side // int(getattr(config, "mm_tokens_per_image", 256) ** 0.5)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, pixel_values):
        n, C, H, W = pixel_values.shape
        P = self.vc.patch_size
        x = pixel_values.reshape(n, C, H // P, P, W // P, P).permute(0, 2, 4, 1, 3, 5).reshape(n, -1, C * P * P)
        x = self.patch_embedding(x) + self.position_embedding.weight
        for layer in self.layers:
            x = layer(x)
        x = self.post_layernorm(x)
        g = x.transpose(1, 2).reshape(n, -1, self.side, self.side)
        g = F.avg_pool2d(g, self.pool, self.pool).flatten(2).transpose(1, 2)          # [n, mm_tokens, H_vision]
        out = self.mm_soft_emb_norm(g) @ self.mm_input_projection_weight
        return out.reshape(-1, out.shape[-1])


class NeuronGemma3MMTextModel(NeuronGemma3Model):
    meta_extra_keys = ("bidir_group_ids",)
    graph_safe = False

    def embed(self, input_ids, inputs_embeds=None, vision_embeddings=None, vision_mask=None):
        """Token embeddings are scaled by sqrt(H); the projected image features are inserted UNSCALED."""
        h = inputs_embeds if inputs_embeds is not None else self.embed_tokens(input_ids)
        h = h * torch.tensor(self.embed_scale, dtype=h.dtype, device=h.device)
        if vision_embeddings is not None and vision_mask is not None:
            h = self.encode_vision_to_input(h, vision_embeddings, vision_mask)
        return h


class NeuronGemma3ForConditionalGeneration(NeuronBaseForImageToText):
    _model_cls = NeuronGemma3MMTextModel
    _vision_cls = NeuronGemma3VisionModel
    text_prefix = "language_model."
    vision_prefix = "vision_tower."
    vision_kwargs = ()

    @classmethod
    def get_config_cls(cls):
        return Gemma3VisionInferenceConfig

    @staticmethod
    def load_hf_model(model_path):
        from transformers import AutoModelForImageTextToText
        return AutoModelForImageTextToText.from_pretrained(model_path)

    convert_hf_to_neuron_state_dict = staticmethod(NeuronGemma3ForCausalLM.convert_hf_to_neuron_state_dict)

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        sd["lm_head.weight"] = sd["embed_tokens.weight"].clone()

    @classmethod
    def get_state_dict(cls, path, config):
        from ...modules.checkpoint import load_state_dict
        sd = {cls._strip(k): v for k, v in load_state_dict(path).items()}
        text = {(k[len(cls.text_prefix):] if k.startswith(cls.text_prefix) else k): v for k, v in sd.items()
                if not k.startswith(cls.vision_prefix) and not k.startswith("multi_modal_projector.")}
        text = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in text.items()}
        text = cls.convert_hf_to_neuron_state_dict(text, config.get_text_config())
        if "lm_head.weight" not in text:
            cls.update_state_dict_for_tied_weights(text)
        out = {cls.text_prefix + k: v for k, v in text.items()}
        vis = {}
        for k, v in sd.items():
            if k.startswith("multi_modal_projector."):
                vis[k[len("multi_modal_projector."):]] = v
                continue
            if not k.startswith(cls.vision_prefix):
                continue
            k = k[len(cls.vision_prefix):].replace("vision_model.", "")
            if k.startswith("head.") or "position_ids" in k:
                continue                                                         # SigLIP pooling head is unused
            k = (k.replace("embeddings.position_embedding.", "position_embedding.").replace("embeddings.patch_embedding.", "patch_embedding.proj.")
                 .replace("encoder.layers.", "layers.").replace(".self_attn.out_proj.", ".self_attn.o_proj."))
            if k == "patch_embedding.proj.weight":
                v = v.reshape(v.shape[0], -1)
            vis[k] = v
        vis = fuse_qkv_and_gate_up(vis, config.vision_config.num_hidden_layers, fuse_mlp=False)
        out.update({cls.vision_prefix + k: v for k, v in vis.items()})
        return out

    def encode_images(self, pixel_values, **kw):
        return self.vision_encoder_model(pixel_values)

    def forward(self, input_ids, attention_mask=None, position_ids=None, seq_ids=None, sampling_params=None, pixel_values=None,
                vision_embeddings=None, vision_mask=None, token_type_ids=None, **kw):
        if input_ids.shape[-1] > 1 and (pixel_values is not None or vision_embeddings is not None or token_type_ids is not None):
            is_img = token_type_ids.bool() if token_type_ids is not None else torch.zeros_like(input_ids, dtype=torch.bool)
            if token_type_ids is None:
                for t in self.image_token_ids():
                    is_img |= input_ids == t
            start = is_img & ~F.pad(is_img, (1, 0))[:, :-1]
            kw["bidir_group_ids"] = torch.where(is_img, torch.cumsum(start.int(), 1) - 1, torch.full_like(input_ids, -1, dtype=torch.int32))
        return super().forward(input_ids, attention_mask, position_ids, seq_ids, sampling_params, pixel_values=pixel_values,
                               vision_embeddings=vision_embeddings, vision_mask=vision_mask, **kw)
