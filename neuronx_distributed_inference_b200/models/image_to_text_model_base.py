"""Image-to-text applications: a vision encoder feeding a causal text decoder.

Role of the reference's ``NeuronBaseForImageToText`` (models/image_to_text_model_base.py:118-773; separate text and vision
model builders, vision embeddings scattered into the token embeddings at ``vision_mask`` positions, M-RoPE position plumbing)
and ``ImageToTextInferenceConfig`` (``text_config`` / ``vision_config`` each with their own NeuronConfig).

On B200 both towers live in one process per GPU: the vision encoder is an eager prefill-like module (``EncoderRunner``), its
output stays on the device and is consumed by the text model's context-encoding runner in the same stream — no host round
trip between the two "sub-models" as in the reference's two-NEFF design."""
from __future__ import annotations

import copy
from typing import List, Optional

import torch

from ..config import InferenceConfig, NeuronConfig
from ..modules.checkpoint import load_sharded
from .application_base import NeuronBaseForCausalLM
from .encoder_base import VISION_ENCODER_MODEL_TAG, EncoderRunner


class ImageToTextInferenceConfig(InferenceConfig):
    """``text_config`` / ``vision_config`` namespaces; ``vision_neuron_config`` optionally overrides the NeuronConfig of the
    vision tower (reference image_to_text_model_base.py:40-115)."""

    def __init__(self, neuron_config=None, vision_neuron_config=None, text_neuron_config=None, **kw):
        self._vision_nc = vision_neuron_config
        super().__init__(text_neuron_config or neuron_config, **kw)

    def __setattr__(self, k, v):
        object.__setattr__(self, k, v) if k.startswith("_") else super().__setattr__(k, v)

    def add_derived_config(self):
        self.num_cores_per_group = 1
        tc = getattr(self, "text_config", None)
        if tc is not None:
            object.__setattr__(tc, "neuron_config", self.neuron_config)
            if getattr(tc, "head_dim", None) is None:
                object.__setattr__(tc, "head_dim", tc.hidden_size // tc.num_attention_heads)
            for k, d in (("hidden_act", "silu"), ("rms_norm_eps", 1e-6), ("pad_token_id", getattr(self, "pad_token_id", 0)),
                         ("num_cores_per_group", 1)):
                if not hasattr(tc, k) or getattr(tc, k) is None:
                    object.__setattr__(tc, k, d)
        vc = getattr(self, "vision_config", None)
        if vc is not None:
            object.__setattr__(vc, "neuron_config", self._vision_nc or self.neuron_config)

    def get_text_config(self):
        return getattr(self, "text_config", None) or self


class NeuronBaseForImageToText(NeuronBaseForCausalLM):
    """Subclasses set ``_model_cls`` (text device model), ``_vision_cls`` (vision tower ``nn.Module(config, device)``), the
    state-dict prefixes and implement ``encode_images`` if the vision call needs more than ``pixel_values``."""

    _vision_cls = None
    text_prefix = "language_model."
    vision_prefix = "visual."

    # ---- construction ---------------------------------------------------------------------------------------------
    def _build_module(self, device):
        cfg = self.config.get_text_config()
        with torch.device(device):
            model = self._model_cls(cfg, device=device)
        return model.eval()

    def _build_vision(self, device):
        with torch.device(device):
            return self._vision_cls(self.config, device=device).eval()

    def checkpoint_loader_fn(self, mmap: bool = False) -> dict:
        sd = super().checkpoint_loader_fn(mmap)
        return self._split_state_dict(sd)

    def _split_state_dict(self, sd: dict) -> dict:
        text, vision = {}, {}
        for k, v in sd.items():
            if k.startswith(self.text_prefix):
                text[k[len(self.text_prefix):]] = v
            elif k.startswith(self.vision_prefix):
                vision[k[len(self.vision_prefix):]] = v
            else:
                text[k] = v
                vision[k] = v
        self._vision_sd = vision
        return text

    @classmethod
    def get_state_dict(cls, model_name_or_path, config):
        # conversion runs on the text part after the split; keep raw names here
        from ..modules.checkpoint import load_state_dict
        sd = load_state_dict(model_name_or_path)
        sd = {cls._strip(k): v for k, v in sd.items()}
        text = {k[len(cls.text_prefix):] if k.startswith(cls.text_prefix) else k: v for k, v in sd.items()
                if not k.startswith(cls.vision_prefix)}
        text = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in text.items()}   # legacy nested layout
        text = cls.convert_hf_to_neuron_state_dict(text, config.get_text_config())
        if getattr(config.get_text_config(), "tie_word_embeddings", False) or getattr(config, "tie_word_embeddings", False):
            if "lm_head.weight" not in text:
                cls.update_state_dict_for_tied_weights(text)
        out = {cls.text_prefix + k: v for k, v in text.items()}
        vis = {k[len(cls.vision_prefix):]: v for k, v in sd.items() if k.startswith(cls.vision_prefix)}
        vis = cls.convert_hf_to_neuron_vision_state_dict(vis, config)
        out.update({cls.vision_prefix + k: v for k, v in vis.items()})
        return out

    @staticmethod
    def convert_hf_to_neuron_vision_state_dict(sd: dict, config) -> dict:
        return sd

    def _post_load(self, model):
        super()._post_load(model)
        self.vision_model = self._build_vision(self.device)
        vsd = getattr(self, "_vision_sd", None)
        if vsd:
            load_sharded(self.vision_model, vsd, self.config.vision_config.neuron_config.torch_dtype, strict=False)
            self._vision_sd = None
        else:   # random weights
            main, self.model = self.model, self.vision_model
            try:
                self.init_random_weights(31)
            finally:
                self.model = main

    def _build_runners(self):
        super()._build_runners()
        vnc = self.config.vision_config.neuron_config
        self.vision_encoder_model = EncoderRunner(VISION_ENCODER_MODEL_TAG, self.vision_model,
                                                  getattr(vnc, "vision_buckets", None), 0, self.device)
        self.models.append(self.vision_encoder_model)

    # ---- forward --------------------------------------------------------------------------------------------------
    def image_token_ids(self) -> List[int]:
        ids = [getattr(self.config, n, None) for n in ("image_token_id", "image_token_index", "video_token_id")]
        return [i for i in ids if i is not None]

    def encode_images(self, pixel_values, **kw) -> torch.Tensor:
        """-> vision embeddings ``[n_image_tokens, H_text]`` in the order the placeholder tokens appear."""
        return self.vision_encoder_model(pixel_values, **kw)

    def get_rotary_position_ids(self, input_ids, attention_mask, **kw) -> Optional[torch.Tensor]:
        return None

    def forward(self, input_ids, attention_mask=None, position_ids=None, seq_ids=None, sampling_params=None,
                pixel_values=None, vision_embeddings=None, vision_mask=None, rotary_position_ids=None, **kw):
        vis_kw = {k: kw.pop(k) for k in list(kw) if k in self.vision_kwargs}
        is_prefill = input_ids.shape[-1] > 1 and (position_ids is None or int(position_ids.reshape(-1)[0]) == 0
                                                  or min(position_ids[:, 0].tolist()) == 0)
        if is_prefill:
            if pixel_values is not None and vision_embeddings is None:
                vision_embeddings = self.encode_images(pixel_values, **vis_kw)
            if vision_embeddings is not None and vision_mask is None:
                vision_mask = torch.zeros_like(input_ids, dtype=torch.bool)
                for t in self.image_token_ids():
                    vision_mask |= input_ids == t
            if rotary_position_ids is None:
                rotary_position_ids = self.get_rotary_position_ids(input_ids, attention_mask, **vis_kw)
                if rotary_position_ids is not None:
                    # decode continues from max(position)+1 on every axis: remember the per-row offset
                    n = (attention_mask.long().sum(-1) if attention_mask is not None
                         else torch.full((input_ids.shape[0],), input_ids.shape[1]))
                    self._rope_delta = (rotary_position_ids.amax(dim=(0, 2)) + 1 - n).view(-1, 1)
                else:
                    self._rope_delta = None
        elif rotary_position_ids is None and getattr(self, "_rope_delta", None) is not None and position_ids is not None:
            p = position_ids.to(self._rope_delta.device) + self._rope_delta[: position_ids.shape[0]]
            rotary_position_ids = p.unsqueeze(0).expand(3, -1, -1)
        return super().forward(input_ids, attention_mask, position_ids, seq_ids, sampling_params,
                               vision_embeddings=vision_embeddings, vision_mask=vision_mask,
                               rotary_position_ids=rotary_position_ids, **kw)

    vision_kwargs = ("image_grid_thw", "image_sizes", "aspect_ratio_ids", "aspect_ratio_mask", "video_grid_thw", "pixel_values_videos")

    def reset(self):
        super().reset()
        self._rope_delta = None
