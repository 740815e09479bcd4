"""Encoder-only applications (vision towers, audio encoders, text encoders of diffusion pipelines).

Role of the reference's ``NeuronEncoderBase`` / ``NeuronEncoderApplication`` (models/encoder_base.py:18-179): a model without
KV cache, one sub-model, inputs bucketed on a single dimension.  On B200 the encoder runs eagerly on the device (prefill-like,
compute bound: the tcgen05 GEMM path); inputs are padded to the bucket, copied through pinned staging, outputs unpadded."""
from __future__ import annotations

import logging
import time
from typing import List, Optional

import torch
import torch.nn as nn

from ..modules.checkpoint import load_sharded
from .application_base import NeuronApplicationBase

logger = logging.getLogger("b200infer")

VISION_ENCODER_MODEL_TAG = "vision_encoder_model"


class NeuronEncoderBase(nn.Module):
    """Device module base: subclasses define ``forward(*tensors)``."""

    def __init__(self, config, device=None):
        super().__init__()
        self.config = config
        self.neuron_config = config.neuron_config
        self.device_ = device


class EncoderRunner:
    """Bucket + pad + run (the encoder counterpart of runtime.runner.SubModelRunner)."""

    def __init__(self, tag: str, module: nn.Module, buckets: Optional[List[int]] = None, pad_dim: int = 0, device=None):
        self.tag, self.module, self.buckets, self.pad_dim, self.device = tag, module, sorted(buckets or []), pad_dim, device
        self.collector = None

    def pick_bucket(self, n: int) -> int:
        for b in self.buckets:
            if b >= n:
                return b
        if self.buckets:
            raise ValueError(f"{self.tag}: input of size {n} exceeds the largest bucket {self.buckets[-1]}")
        return n

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, *args, **kw):
        t0 = time.perf_counter()
        n = x.shape[self.pad_dim]
        b = self.pick_bucket(n)
        if b != n:
            pad = list(x.shape)
            pad[self.pad_dim] = b - n
            x = torch.cat([x, x.new_zeros(pad)], self.pad_dim)
        x = x.to(self.device, non_blocking=True)
        args = [a.to(self.device, non_blocking=True) if torch.is_tensor(a) else a for a in args]
        kw = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in kw.items()}
        out = self.module(x, *args, n_valid=n, **kw) if getattr(self.module, "takes_n_valid", False) else self.module(x, *args, **kw)
        if self.collector is not None:
            if self.device is not None and self.device.type == "cuda":
                torch.cuda.synchronize()
            self.collector.add(self.tag, time.perf_counter() - t0)
        return out

    def reset(self):
        pass

    def warmup(self):
        pass


class NeuronEncoderApplication(NeuronApplicationBase):
    """Application wrapper for a single encoder module (``_model_cls`` builds it)."""

    encoder_tag = VISION_ENCODER_MODEL_TAG

    def _build_runners(self):
        nc = self.neuron_config
        self.encoder_model = EncoderRunner(self.encoder_tag, self.model, getattr(nc, "buckets", None), 0, self.device)
        self.models = [self.encoder_model]

    def forward(self, *args, **kw):
        return self.encoder_model(*args, **kw)
