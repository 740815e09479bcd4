"""FLUX application: four sub-applications (CLIP, T5, MMDiT backbone, VAE decoder) behind one pipeline object.

reference: models/diffusers/flux/application.py (``NeuronFluxApplication``; context-parallel or CFG-parallel with dp=2 :33-65).
On B200 every sub-model is an eager encoder-style module on the same device; the backbone is tensor parallel over heads.  With
``world = 2 x tp`` the two replicas of the TP group either split the IMAGE TOKENS (``context_parallel_enabled``: queries local, keys /
values all-gathered per attention) or the two branches of true classifier-free guidance (``cfg_parallel_enabled``)."""
from __future__ import annotations

import json
import os
from typing import Optional

import torch

from ....config import InferenceConfig, NeuronConfig
from ....modules.checkpoint import load_sharded, load_state_dict
from ....parallel import state as pstate
from ...encoder_base import EncoderRunner
from .clip.modeling_clip import NeuronClipTextModel, convert_clip_state_dict
from .modeling_flux import FluxBackboneInferenceConfig, NeuronFluxTransformer2DModel, convert_diffusers_flux_state_dict
from .pipeline import FlowMatchEulerScheduler, NeuronFluxControlPipeline, NeuronFluxFillPipeline, NeuronFluxPipeline
from .t5.modeling_t5 import NeuronT5EncoderModel, convert_t5_state_dict
from .vae.modeling_vae import NeuronVAEDecoder, NeuronVAEEncoder, convert_vae_decoder_state_dict, convert_vae_encoder_state_dict


def get_flux_parallelism_config(backbone_tp_degree: int, context_parallel_enabled: bool = False, cfg_parallel_enabled: bool = False) -> int:
    """World size for a backbone TP degree: doubled when one of the two (mutually exclusive) dp=2 modes is on (reference :33-65)."""
    if context_parallel_enabled and cfg_parallel_enabled:
        raise ValueError("context_parallel_enabled and cfg_parallel_enabled are mutually exclusive")
    return backbone_tp_degree * (2 if (context_parallel_enabled or cfg_parallel_enabled) else 1)


def _ns(neuron_config, d: dict) -> InferenceConfig:
    ns = InferenceConfig.__new__(InferenceConfig)
    object.__setattr__(ns, "neuron_config", neuron_config)
    for k, v in d.items():
        object.__setattr__(ns, k, v)
    return ns


class NeuronFluxApplication:
    """``model_path``: a diffusers FLUX directory (``transformer/``, ``text_encoder/``, ``text_encoder_2/``, ``vae/``), or
    ``None`` with explicit config dicts for random-weight benchmarking."""

    def __init__(self, model_path: Optional[str], neuron_config: Optional[NeuronConfig] = None, backbone_config: Optional[dict] = None,
                 clip_config: Optional[dict] = None, t5_config: Optional[dict] = None, vae_config: Optional[dict] = None,
                 height: int = 1024, width: int = 1024, task: str = "text-to-image", context_parallel_enabled: bool = False,
                 cfg_parallel_enabled: bool = False):
        """``task``: ``text-to-image`` (FLUX.1 dev / schnell), ``control`` (Canny / Depth dev: backbone in_channels 128) or ``fill``
        (Fill dev: in_channels 384); the image-conditioned tasks also load the VAE encoder."""
        if task not in ("text-to-image", "control", "fill"):
            raise ValueError(f"unknown FLUX task {task}")
        self.task = task
        self.context_parallel_enabled, self.cfg_parallel_enabled = context_parallel_enabled, cfg_parallel_enabled
        get_flux_parallelism_config(1, context_parallel_enabled, cfg_parallel_enabled)          # validates exclusivity
        self.model_path = model_path
        self.neuron_config = neuron_config or NeuronConfig(batch_size=1, torch_dtype="bfloat16")
        self.height, self.width = height, width

        def cfg(sub, given):
            if given is not None:
                return given
            with open(os.path.join(model_path, sub, "config.json")) as f:
                return json.load(f)
        nc = self.neuron_config
        bc = cfg("transformer", backbone_config)
        self.backbone_config = FluxBackboneInferenceConfig(nc, load_config=lambda c: [setattr(c, k, v) for k, v in bc.items()
                                                                                      if not k.startswith("_")])
        self.clip_config = _ns(nc, cfg("text_encoder", clip_config))
        self.t5_config = _ns(nc, cfg("text_encoder_2", t5_config))
        self.vae_config = _ns(nc, cfg("vae", vae_config))
        self.pipe = None

    def compile(self, path: str, **kw):
        os.makedirs(path, exist_ok=True)
        self.backbone_config.save(os.path.join(path, "transformer"))

    def load(self, path: Optional[str] = None, random_weights: bool = False, seed: int = 0):
        nc = self.neuron_config
        dp2 = self.context_parallel_enabled or self.cfg_parallel_enabled
        if nc.on_cpu or not torch.cuda.is_available():
            dev = torch.device("cpu")
            if nc.tp_degree > 1 or dp2:
                pstate.init_distributed("gloo")
        else:
            pstate.init_distributed("nccl")
            dev = torch.device("cuda", torch.cuda.current_device())
        if not pstate.model_parallel_is_initialized():
            pstate.initialize_model_parallel(tensor_model_parallel_size=nc.tp_degree)
        self.device = dev
        dt = nc.torch_dtype
        with torch.device(dev):
            self.transformer = NeuronFluxTransformer2DModel(self.backbone_config, dev).eval()
            self.clip = NeuronClipTextModel(self.clip_config, dev).eval()
            self.t5 = NeuronT5EncoderModel(self.t5_config, dev).eval()
            self.vae = NeuronVAEDecoder(self.vae_config, dev).eval()
            self.vae_encoder = NeuronVAEEncoder(self.vae_config, dev).eval() if self.task != "text-to-image" else None
        if random_weights:
            g = torch.Generator(device=dev).manual_seed(seed)
            for m in filter(None, (self.transformer, self.clip, self.t5, self.vae, self.vae_encoder)):
                for n, p in m.named_parameters():
                    if p.dim() == 1 and "norm" in n and "bias" not in n:
                        p.fill_(1.0)
                    elif p.dim() == 1:
                        p.zero_()
                    else:
                        p.copy_(torch.randn(p.shape, device=dev, generator=g, dtype=torch.float32) * 0.02)
        else:
            mp = self.model_path
            load_sharded(self.transformer, convert_diffusers_flux_state_dict(load_state_dict(os.path.join(mp, "transformer")),
                                                                             self.backbone_config), dt, strict=False)
            load_sharded(self.clip, convert_clip_state_dict(load_state_dict(os.path.join(mp, "text_encoder")), self.clip_config), dt, strict=False)
            load_sharded(self.t5, convert_t5_state_dict(load_state_dict(os.path.join(mp, "text_encoder_2")), self.t5_config), dt, strict=False)
            load_sharded(self.vae, convert_vae_decoder_state_dict(load_state_dict(os.path.join(mp, "vae"))), dt, strict=False)
            if self.vae_encoder is not None:
                load_sharded(self.vae_encoder, convert_vae_encoder_state_dict(load_state_dict(os.path.join(mp, "vae"))), dt, strict=False)
        self.models = [EncoderRunner("clip_text_encoder", self.clip, None, 0, dev), EncoderRunner("t5_text_encoder", self.t5, None, 0, dev),
                       EncoderRunner("flux_backbone", self.transformer, None, 0, dev), EncoderRunner("vae_decoder", self.vae, None, 0, dev)]
        pipe_cls = {"text-to-image": NeuronFluxPipeline, "control": NeuronFluxControlPipeline, "fill": NeuronFluxFillPipeline}[self.task]
        extra = {}
        if self.vae_encoder is not None:
            self.models.append(EncoderRunner("vae_encoder", self.vae_encoder, None, 0, dev))
            extra["vae_encoder"] = self.models[-1]
        if dp2:
            rep = pstate.get_data_parallel_group()
            if rep.size != 2:
                raise RuntimeError(f"context / CFG parallel need world = 2 x tp_degree ({2 * nc.tp_degree}); got {rep.size} replica(s)")
            if self.context_parallel_enabled:
                self.transformer.cp_group = rep
            else:
                extra["cfg_group"] = rep
        self.pipe = pipe_cls(self.models[2], self.models[0], self.models[1], self.models[3], FlowMatchEulerScheduler(),
                             2 ** (len(self.vae_config.block_out_channels) - 1), self.vae_config.latent_channels, dev, dt,
                             self.backbone_config.guidance_embeds, **extra)
        return self

    def __call__(self, clip_input_ids, t5_input_ids, **kw):
        kw.setdefault("height", self.height)
        kw.setdefault("width", self.width)
        return self.pipe(clip_input_ids, t5_input_ids, **kw)
