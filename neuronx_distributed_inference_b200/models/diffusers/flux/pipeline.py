"""FLUX text-to-image pipeline (reference models/diffusers/flux/pipeline.py subclasses diffusers' FluxPipeline; diffusers is
not a dependency here, so the pieces it provides are implemented: latent packing, position ids, the flow-matching Euler
schedule with resolution-dependent time shift, classifier-free-guidance-free *guidance distillation* input)."""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch


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
                 vae_scale_factor: int = 8, latent_channels: int = 16, device=None, dtype=torch.bfloat16, guidance_embeds=True):
        self.transformer, self.clip, self.t5, self.vae = transformer, clip, t5, vae_decoder
        self.scheduler = scheduler or FlowMatchEulerScheduler()
        self.vsf, self.lc, self.device, self.dtype, self.guidance_embeds = vae_scale_factor, latent_channels, device, dtype, guidance_embeds

    @torch.no_grad()
    def __call__(self, clip_input_ids, t5_input_ids, height=1024, width=1024, num_inference_steps=28, guidance_scale=3.5,
                 generator=None, latents=None, output_type="pt"):
        dev = self.device
        B = clip_input_ids.shape[0]
        _, pooled = self.clip(clip_input_ids.to(dev))
        prompt = self.t5(t5_input_ids.to(dev))
        H, W = 2 * (height // (self.vsf * 2)), 2 * (width // (self.vsf * 2))
        if latents is None:
            latents = torch.randn(B, self.lc, H, W, generator=generator, dtype=torch.float32).to(dev, self.dtype)
        x = pack_latents(latents)
        img_ids = latent_image_ids(H // 2, W // 2, dev)
        txt_ids = torch.zeros(prompt.shape[1], 3, device=dev)
        ts = self.scheduler.set_timesteps(num_inference_steps, calculate_shift(x.shape[1]))
        g = torch.full((B,), guidance_scale, device=dev, dtype=torch.float32) if self.guidance_embeds else None
        for i, t in enumerate(ts):
            tt = (t / 1000).expand(B).to(dev)
            v = self.transformer(x, prompt, pooled, tt, img_ids, txt_ids, g)
            x = self.scheduler.step(v.to(x.dtype), i, x)
        img = self.vae(unpack_latents(x, H, W))
        if output_type == "latent":
            return unpack_latents(x, H, W)
        return (img.float() / 2 + 0.5).clamp(0, 1)
