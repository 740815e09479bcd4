"""Llama-4 image-text-to-text (reference models/llama4/modeling_llama4.py:1-451): vision tower + the MoE text decoder."""
from __future__ import annotations

from ..image_to_text_model_base import ImageToTextInferenceConfig, NeuronBaseForImageToText
from ...config import MoENeuronConfig
from .modeling_llama4_text import NeuronLlama4TextForCausalLM, NeuronLlama4TextModel
from .modeling_llama4_vision import NeuronLlama4VisionModel


class Llama4MultimodalInferenceConfig(ImageToTextInferenceConfig):
    def add_derived_config(self):
        super().add_derived_config()
        tc = self.text_config
        n = tc.num_hidden_layers
        if not getattr(tc, "no_rope_layers", None):
            step = getattr(tc, "no_rope_layer_interval", 4)
            object.__setattr__(tc, "no_rope_layers", [int((i + 1) % step != 0) for i in range(n)])
        if not getattr(tc, "moe_layers", None):
            step = getattr(tc, "interleave_moe_layer_step", 1)
            object.__setattr__(tc, "moe_layers", list(range(step - 1, n, step)))

    @classmethod
    def get_neuron_config_cls(cls):
        return MoENeuronConfig


class NeuronLlama4ForCausalLM(NeuronBaseForImageToText):
    _model_cls = NeuronLlama4TextModel
    _vision_cls = NeuronLlama4VisionModel
    text_prefix = "language_model."
    vision_prefix = "vision_model."
    _STATE_DICT_MODEL_PREFIX = ""

    @classmethod
    def get_config_cls(cls):
        return Llama4MultimodalInferenceConfig

    @staticmethod
    def load_hf_model(model_path):
        from transformers import AutoModelForImageTextToText
        return AutoModelForImageTextToText.from_pretrained(model_path)

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        return NeuronLlama4TextForCausalLM.convert_hf_to_neuron_state_dict(sd, config)

    @classmethod
    def get_state_dict(cls, path, config):
        from ...modules.checkpoint import load_state_dict
        from ..state_dict_utils import fuse_qkv_and_gate_up
        sd = load_state_dict(path)
        sd = {(k[len("model."):] if k.startswith("model.") and not k.startswith("model.layers") else k): v for k, v in sd.items()}
        text = {k: v for k, v in sd.items() if k.startswith(cls.text_prefix)}
        text = cls.convert_hf_to_neuron_state_dict(text, config.get_text_config())
        out = {cls.text_prefix + k: v for k, v in text.items()}
        vis = {}
        for k, v in sd.items():
            if k.startswith("multi_modal_projector.linear_1."):
                vis["projector." + k.rsplit(".", 1)[1]] = v
            elif k.startswith(cls.vision_prefix):
                k = k[len(cls.vision_prefix):]
                k = (k.replace("model.layers.", "layers.").replace("vision_adapter.mlp.fc1.", "adapter_fc1.")
                     .replace("vision_adapter.mlp.fc2.", "adapter_fc2.").replace("patch_embedding.linear.", "patch_embedding.proj."))
                vis[k] = v
        vis = fuse_qkv_and_gate_up(vis, config.vision_config.num_hidden_layers, fuse_mlp=False)
        out.update({cls.vision_prefix + k: v for k, v in vis.items()})
        return out

    def encode_images(self, pixel_values, **kw):
        return self.vision_encoder_model(pixel_values)
