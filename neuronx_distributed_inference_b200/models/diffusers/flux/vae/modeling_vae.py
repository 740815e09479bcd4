"""VAE decoder of FLUX (AutoencoderKL decoder half; reference models/diffusers/flux/vae/modeling_vae.py).  Convolutions go
through cuDNN (library ops, not a hot path of the benchmarked models); group-norm + SiLU fused by torch."""
from __future__ import annotations

import torch
import torch.nn as nn


class ResnetBlock(nn.Module):
    def __init__(self, cin, cout, groups, dtype, device):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6, dtype=dtype, device=device)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, dtype=dtype, device=device)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6, dtype=dtype, device=device)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, dtype=dtype, device=device)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1, dtype=dtype, device=device) if cin != cout else None

    def forward(self, x):
        h = self.conv1(nn.functional.silu(self.norm1(x)))
        h = self.conv2(nn.functional.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class VaeAttention(nn.Module):
    def __init__(self, ch, groups, dtype, device):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6, dtype=dtype, device=device)
        self.to_q, self.to_k, self.to_v, self.to_out = (nn.Linear(ch, ch, dtype=dtype, device=device) for _ in range(4))

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        a = torch.softmax(q.float() @ k.float().transpose(1, 2) * C ** -0.5, -1).to(v.dtype) @ v
        return x + self.to_out(a).transpose(1, 2).reshape(B, C, H, W)


class NeuronVAEDecoder(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        c, dt = config, config.neuron_config.torch_dtype
        self.config = config
        chans = list(c.block_out_channels)
        g = getattr(c, "norm_num_groups", 32)
        top = chans[-1]
        self.conv_in = nn.Conv2d(c.latent_channels, top, 3, padding=1, dtype=dt, device=device)
        self.mid_res1 = ResnetBlock(top, top, g, dt, device)
        self.mid_attn = VaeAttention(top, g, dt, device)
        self.mid_res2 = ResnetBlock(top, top, g, dt, device)
        self.up_blocks = nn.ModuleList()
        cin = top
        for i, cout in enumerate(reversed(chans)):
            blk = nn.Module()
            blk.resnets = nn.ModuleList([ResnetBlock(cin if j == 0 else cout, cout, g, dt, device)
                                         for j in range(getattr(c, "layers_per_block", 2) + 1)])
            blk.upsample = nn.Conv2d(cout, cout, 3, padding=1, dtype=dt, device=device) if i < len(chans) - 1 else None
            self.up_blocks.append(blk)
            cin = cout
        self.conv_norm_out = nn.GroupNorm(g, chans[0], eps=1e-6, dtype=dt, device=device)
        self.conv_out = nn.Conv2d(chans[0], c.out_channels, 3, padding=1, dtype=dt, device=device)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, latents):
        c = self.config
        z = latents.to(self.conv_in.weight.dtype) / getattr(c, "scaling_factor", 1.0) + getattr(c, "shift_factor", 0.0)
        h = self.mid_res2(self.mid_attn(self.mid_res1(self.conv_in(z))))
        for blk in self.up_blocks:
            for r in blk.resnets:
                h = r(h)
            if blk.upsample is not None:
                h = blk.upsample(nn.functional.interpolate(h, scale_factor=2.0, mode="nearest"))
        return self.conv_out(nn.functional.silu(self.conv_norm_out(h)))


def convert_vae_decoder_state_dict(sd: dict) -> dict:
    out = {}
    for k, v in sd.items():
        if not k.startswith("decoder."):
            continue
        k = k[len("decoder."):]
        k = (k.replace("mid_block.resnets.0.", "mid_res1.").replace("mid_block.resnets.1.", "mid_res2.")
             .replace("mid_block.attentions.0.", "mid_attn.").replace(".upsamplers.0.conv.", ".upsample.").replace(".to_out.0.", ".to_out."))
        out[k] = v
    return out


class NeuronVAEEncoder(nn.Module):
    """AutoencoderKL encoder half (image -> latent mean), needed by the image-conditioned FLUX pipelines (Control / Fill): conv_in,
    per resolution ``layers_per_block`` ResNet blocks + a stride-2 convolution with (0,1,0,1) padding, the mid block, GroupNorm-SiLU-
    conv_out to ``2 * latent_channels`` (mean | logvar), optional 1x1 ``quant_conv``.  ``forward`` returns the scaled MODE of the
    posterior: ``(mean - shift_factor) * scaling_factor``."""

    def __init__(self, config, device=None):
        super().__init__()
        c, dt = config, config.neuron_config.torch_dtype
        self.config = config
        chans = list(c.block_out_channels)
        g = getattr(c, "norm_num_groups", 32)
        self.conv_in = nn.Conv2d(getattr(c, "in_channels", getattr(c, "out_channels", 3)), chans[0], 3, padding=1, dtype=dt, device=device)
        self.down_blocks = nn.ModuleList()
        cin = chans[0]
        for i, cout in enumerate(chans):
            blk = nn.Module()
            blk.resnets = nn.ModuleList([ResnetBlock(cin if j == 0 else cout, cout, g, dt, device)
                                         for j in range(getattr(c, "layers_per_block", 2))])
            blk.downsample = nn.Conv2d(cout, cout, 3, stride=2, padding=0, dtype=dt, device=device) if i < len(chans) - 1 else None
            self.down_blocks.append(blk)
            cin = cout
        top = chans[-1]
        self.mid_res1 = ResnetBlock(top, top, g, dt, device)
        self.mid_attn = VaeAttention(top, g, dt, device)
        self.mid_res2 = ResnetBlock(top, top, g, dt, device)
        self.conv_norm_out = nn.GroupNorm(g, top, eps=1e-6, dtype=dt, device=device)
        self.conv_out = nn.Conv2d(top, 2 * c.latent_channels, 3, padding=1, dtype=dt, device=device)
        self.quant_conv = nn.Conv2d(2 * c.latent_channels, 2 * c.latent_channels, 1, dtype=dt, device=device) \
            if getattr(c, "use_quant_conv", False) else None
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, image):
        c = self.config
        h = self.conv_in(image.to(self.conv_in.weight.dtype))
        for blk in self.down_blocks:
            for r in blk.resnets:
                h = r(h)
            if blk.downsample is not None:
                h = blk.downsample(nn.functional.pad(h, (0, 1, 0, 1)))
        h = self.mid_res2(self.mid_attn(self.mid_res1(h)))
        moments = self.conv_out(nn.functional.silu(self.conv_norm_out(h)))
        if self.quant_conv is not None:
            moments = self.quant_conv(moments)
        mean = moments[:, : c.latent_channels]
        return (mean - getattr(c, "shift_factor", 0.0)) * getattr(c, "scaling_factor", 1.0)


def convert_vae_encoder_state_dict(sd: dict) -> dict:
    out = {}
    for k, v in sd.items():
        if k.startswith("quant_conv."):
            out[k] = v
            continue
        if not k.startswith("encoder."):
            continue
        k = k[len("encoder."):]
        k = (k.replace("mid_block.resnets.0.", "mid_res1.").replace("mid_block.resnets.1.", "mid_res2.")
             .replace("mid_block.attentions.0.", "mid_attn.").replace(".downsamplers.0.conv.", ".downsample.").replace(".to_out.0.", ".to_out."))
        out[k] = v
    return out
