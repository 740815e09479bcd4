"""FLUX text-to-image pipeline (reference models/diffusers/flux/pipeline.py subclasses diffusers' FluxPipeline; diffusers is
not a dependency here, so the pieces it provides are implemented: latent packing, position ids, the flow-matching Euler
schedule with resolution-dependent time shift, classifier-free-guidance-free *guidance distillation* input)."""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch

from ....parallel import mappings


def calculate_shift(seq_len, base_len=256, max_len=4096, base_shift=0.5, max_shift=1.15):
    m = (max_shift - base_shift) / (max_len - base_len)
    return seq_len * m + (base_shift - m * base_len)


class FlowMatchEulerScheduler:
    def __init__(self, num_train_timesteps=1000, use_dynamic_shifting=True, shift=3.0):
        self.n_train, self.dynamic, self.shift = num_train_timesteps, use_dynamic_shifting, shift
        self.sigmas = None

    def set_timesteps(self, n: int, mu: Optional[float] = None):
        s = np.linspace(1.0, 1.0 / n, n)
        if self.dynamic and mu is not None:
            s = math.exp(mu) / (math.exp(mu) + (1 / s - 1))
        else:
            s = self.shift * s / (1 + (self.shift - 1) * s)
        self.sigmas = torch.tensor(np.concatenate([s, [0.0]]), dtype=torch.float32)
        self.timesteps = self.sigmas[:-1] * self.n_train
        return self.timesteps

    def step(self, model_output, i: int, sample):
        return sample + (self.sigmas[i + 1] - self.sigmas[i]).to(sample.device, sample.dtype) * model_output


def pack_latents(lat):
    B, C, H, W = lat.shape
    return lat.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 2) * (W // 2), C * 4)


def unpack_latents(x, H, W):
    B, N, C4 = x.shape
    return x.view(B, H // 2, W // 2, C4 // 4, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, C4 // 4, H, W)


def latent_image_ids(H2, W2, device):
    ids = torch.zeros(H2, W2, 3, device=device)
    ids[..., 1] += torch.arange(H2, device=device)[:, None]
    ids[..., 2] += torch.arange(W2, device=device)[None, :]
    return ids.view(H2 * W2, 3)


class NeuronFluxPipeline:
    def __init__(self, transformer: Callable, clip: Callable, t5: Callable, vae_decoder: Callable, scheduler=None,
                 vae_scale_factor: int = 8, latent_channels: int = 16, device=None, dtype=torch.bfloat16, guidance_embeds=True,
                 cfg_group=None):
        self.cfg_group = cfg_group
        self.transformer, self.clip, self.t5, self.vae = transformer, clip, t5, vae_decoder
        self.scheduler = scheduler or FlowMatchEulerScheduler()
        self.vsf, self.lc, self.device, self.dtype, self.guidance_embeds = vae_scale_factor, latent_channels, device, dtype, guidance_embeds

    @torch.no_grad()
    def __call__(self, clip_input_ids, t5_input_ids, height=1024, width=1024, num_inference_steps=28, guidance_scale=3.5,
                 generator=None, latents=None, output_type="pt", negative_clip_input_ids=None, negative_t5_input_ids=None,
                 true_cfg_scale: float = 1.0, **cond_kw):
        """``negative_*`` + ``true_cfg_scale > 1``: real classifier-free guidance, ``v = v_neg + s * (v_pos - v_neg)`` (two backbone
        evaluations per step; with ``cfg_group`` of two replicas each computes one of them — the reference's CFG-parallel dp=2)."""
        dev = self.device
        B = clip_input_ids.shape[0]
        _, pooled = self.clip(clip_input_ids.to(dev))
        prompt = self.t5(t5_input_ids.to(dev))
        true_cfg = true_cfg_scale > 1.0 and negative_t5_input_ids is not None
        if true_cfg:
            _, neg_pooled = self.clip((negative_clip_input_ids if negative_clip_input_ids is not None else clip_input_ids).to(dev))
            neg_prompt = self.t5(negative_t5_input_ids.to(dev))
        H, W = 2 * (height // (self.vsf * 2)), 2 * (width // (self.vsf * 2))
        if latents is None:
            latents = torch.randn(B, self.lc, H, W, generator=generator, dtype=torch.float32).to(dev, self.dtype)
        x = pack_latents(latents)
        img_ids = latent_image_ids(H // 2, W // 2, dev)
        txt_ids = torch.zeros(prompt.shape[1], 3, device=dev)
        ts = self.scheduler.set_timesteps(num_inference_steps, calculate_shift(x.shape[1]))
        g = torch.full((B,), guidance_scale, device=dev, dtype=torch.float32) if self.guidance_embeds else None
        cond = self.conditioning(B, H, W, **cond_kw)            # None for text-to-image; extra channels for Control / Fill
        for i, t in enumerate(ts):
            tt = (t / 1000).expand(B).to(dev)
            xin = x if cond is None else torch.cat([x, cond.to(x.dtype)], 2)
            if not true_cfg:
                v = self.transformer(xin, prompt, pooled, tt, img_ids, txt_ids, g)
            elif self.cfg_group is not None and self.cfg_group.size == 2:
                mine = (prompt, pooled) if self.cfg_group.rank == 0 else (neg_prompt, neg_pooled)
                ntxt = torch.zeros(mine[0].shape[1], 3, device=dev)
                both = mappings.all_gather(self.transformer(xin, mine[0], mine[1], tt, img_ids, ntxt, g).unsqueeze(0).contiguous(), 0, self.cfg_group)
                v = both[1] + true_cfg_scale * (both[0] - both[1])
            else:
                vp = self.transformer(xin, prompt, pooled, tt, img_ids, txt_ids, g)
                vn = self.transformer(xin, neg_prompt, neg_pooled, tt, img_ids, torch.zeros(neg_prompt.shape[1], 3, device=dev), g)
                v = vn + true_cfg_scale * (vp - vn)
            x = self.scheduler.step(v.to(x.dtype), i, x)
        img = self.vae(unpack_latents(x, H, W))
        if output_type == "latent":
            return unpack_latents(x, H, W)
        return (img.float() / 2 + 0.5).clamp(0, 1)

    def conditioning(self, B, H, W, **kw):
        """Packed per-token conditioning concatenated to the latent tokens along the channel axis at every step."""
        if kw:
            raise TypeError(f"unexpected arguments {sorted(kw)} for the text-to-image pipeline")
        return None


def _prep_image(img: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """[B,3,h,w] in [0,1] (or already [-1,1] when it has negative values) -> resized to the target size, in [-1,1]."""
    img = img.float()
    if img.shape[-2:] != (height, width):
        img = torch.nn.functional.interpolate(img, size=(height, width), mode="bilinear", align_corners=False)
    return img if img.min() < 0 else img * 2 - 1


class NeuronFluxControlPipeline(NeuronFluxPipeline):
    """FLUX.1 Canny / Depth (reference pipeline.py NeuronFluxControlPipeline): the control image is VAE-encoded, packed like the
    latents and concatenated to them channel-wise, so the backbone runs with ``in_channels = 2 * 64``."""

    def __init__(self, *a, vae_encoder: Callable = None, **kw):
        super().__init__(*a, **kw)
        self.vae_encoder = vae_encoder

    def conditioning(self, B, H, W, control_image=None, **kw):
        if control_image is None:
            raise ValueError("control_image is required")
        if kw:
            raise TypeError(f"unexpected arguments {sorted(kw)}")
        img = _prep_image(control_image, H * self.vsf, W * self.vsf).to(self.device, self.dtype)
        lat = self.vae_encoder(img)
        if lat.shape[0] != B:
            lat = lat.expand(B, -1, -1, -1) if lat.shape[0] == 1 else lat.repeat_interleave(B // lat.shape[0], 0)
        return pack_latents(lat)


class NeuronFluxFillPipeline(NeuronFluxPipeline):
    """FLUX.1 Fill (in/out-painting; reference pipeline.py NeuronFluxFillPipeline): per token the backbone also sees the VAE latents
    of the image with the hole blanked out (64 channels) and the binary mask of its 8x8-pixel x 2x2-latent footprint (256 channels):
    ``in_channels = 64 + 64 + 256 = 384``."""

    def __init__(self, *a, vae_encoder: Callable = None, **kw):
        super().__init__(*a, **kw)
        self.vae_encoder = vae_encoder

    def conditioning(self, B, H, W, image=None, mask_image=None, **kw):
        if image is None or mask_image is None:
            raise ValueError("image and mask_image are required")
        if kw:
            raise TypeError(f"unexpected arguments {sorted(kw)}")
        f = self.vsf
        img = _prep_image(image, H * f, W * f).to(self.device, self.dtype)
        m = mask_image.float()
        if m.dim() == 3:
            m = m.unsqueeze(1)
        if m.shape[-2:] != (H * f, W * f):
            m = torch.nn.functional.interpolate(m, size=(H * f, W * f), mode="nearest")
        m = (m > 0.5).to(self.device, self.dtype)                              # 1 = repaint
        lat = self.vae_encoder(img * (1 - m))
        mk = m[:, 0].view(-1, H, f, W, f).permute(0, 2, 4, 1, 3).reshape(-1, f * f, H, W)     # pixel footprint of every latent position
        out = torch.cat([pack_latents(lat), pack_latents(mk)], 2)
        return out.expand(B, -1, -1) if out.shape[0] == 1 and B > 1 else out
