"""On-device sampling (reference modules/generation/sampling.py:24-579).

``sampling_params`` is a ``[B,3]`` fp32 tensor ``(top_k, top_p, temperature)`` (defaults (1,1,1));
``temperature==0`` or ``top_k==1`` means greedy; ``top_k in {-1,0}`` means "all of global_topk".
With tensor-parallel (vocab-sharded) logits every rank reduces its shard first — local arg-max or
local top-``global_topk`` — and only ``k * 8`` bytes per row cross NVLink before the final choice,
like the reference's staged distributed top-k (sampling.py:285-326) but in one stage: the NVSwitch
makes every peer equidistant, so there is no topology to stage over.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..parallel import mappings
from ..parallel.state import Group, get_tensor_model_parallel_group


def prepare_sampling_params(batch_size: int, top_k=1, top_p=1.0, temperature=1.0) -> torch.Tensor:
    """Broadcast scalars / per-row lists to ``[B,3]`` (reference sampling.py:183-198)."""
    def col(v):
        t = torch.as_tensor(v, dtype=torch.float32).flatten()
        if t.numel() == 1:
            t = t.expand(batch_size)
        assert t.numel() == batch_size, "sampling params must be scalar or one per batch row"
        return t
    return torch.stack([col(top_k), col(top_p), col(temperature)], dim=1).contiguous()


def validate_sampling_params(params: torch.Tensor, on_device_sampling_config) -> None:
    """reference sampling.py:97-157."""
    if params.dim() != 2 or params.shape[1] != 3:
        raise ValueError(f"sampling_params must be [batch,3], got {tuple(params.shape)}")
    top_k, top_p, temp = params[:, 0], params[:, 1], params[:, 2]
    if not torch.equal(top_k, top_k.round()):
        raise ValueError("top_k must be integral")
    gk = on_device_sampling_config.global_topk if on_device_sampling_config else 256
    if ((top_k < -1) | (top_k > gk)).any():
        raise ValueError(f"top_k must be -1, 0 or within [1, global_topk={gk}]")
    if ((top_p <= 0) | (top_p > 1)).any():
        raise ValueError("top_p must be in (0, 1]")
    if (temp < 0).any():
        raise ValueError("temperature must be >= 0")


def infer_sampling_params(params: torch.Tensor) -> torch.Tensor:
    """Special values imply the other columns (reference sampling.py:160-181): ``temperature == 0`` means greedy, so that row's top_k
    and top_p are reset to 1 (and the temperature to 1 so that no division by zero can occur downstream)."""
    params = params.clone()
    greedy = params[:, 2] == 0
    params[greedy, 0] = 1
    params[greedy, 1] = 1
    params[greedy, 2] = 1
    return params


def rand_like(t: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Uniform [0,1) noise shaped like ``t`` on its device (the reference needs an XLA ``Rng`` custom call for this, :88-94)."""
    return torch.rand(t.shape, dtype=torch.float32, device=t.device, generator=generator)


def mask_padded_logits(logits: torch.Tensor, rank: int, world: int, pad_size: int) -> torch.Tensor:
    """Vocab padded to a multiple of tp: the pad columns (all at the end of the last rank's
    shard) must never win (reference sampling.py:24-47)."""
    if pad_size == 0 or rank != world - 1:
        return logits
    logits = logits.clone()
    logits[..., logits.shape[-1] - pad_size:] = torch.finfo(logits.dtype).min
    return logits


class Sampler(nn.Module):
    def __init__(self, neuron_config, tp_group: Optional[Group] = None, vocab_shard: bool = True):
        super().__init__()
        self.cfg = neuron_config.on_device_sampling_config
        self.tp_group = tp_group or get_tensor_model_parallel_group()
        self.vocab_shard = vocab_shard and self.tp_group.size > 1
        self.global_topk = self.cfg.global_topk if self.cfg else 256
        self.do_sample = bool(self.cfg and self.cfg.do_sample)
        self.dynamic = bool(self.cfg and self.cfg.dynamic)
        self.deterministic = bool(self.cfg and self.cfg.deterministic)
        self._gen = None

    def _rand(self, B, device):
        if self.deterministic:
            return torch.full((B,), 0.5, device=device)
        if device.type == "cuda":
            # The device's DEFAULT generator: it is the only one torch registers with a CUDA graph under capture (a private
            # torch.Generator raises "CUDA generator not in capture mode" inside the decode / prefill graphs), and its philox
            # offset advances per replay, so graph replays draw fresh numbers.  Seeded once per sampler from the config.
            if self._gen is None:
                self._gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
                if not torch.cuda.is_current_stream_capturing():
                    self._gen.manual_seed(int(self.cfg.seed) if self.cfg else 0)
            return torch.rand(B, device=device)
        if self._gen is None or self._gen.device != device:
            self._gen = torch.Generator(device=device)
            self._gen.manual_seed(int(self.cfg.seed) if self.cfg else 0)
        return torch.rand(B, device=device, generator=self._gen)

    def greedy(self, logits: torch.Tensor) -> torch.Tensor:
        """logits [B, V_local] -> global token ids [B]."""
        g = self.tp_group
        if not self.vocab_shard:
            return ops.argmax(logits)
        tok = ops.argmax_sharded(logits, g)                       # one kernel: local arg-max + NVLink LL exchange
        if tok is not None:
            return tok
        idx = ops.argmax(logits)                                  # local arg-max kernel
        val = logits.gather(1, idx.view(-1, 1)).view(-1).float()
        idx = idx + g.rank * logits.shape[-1]
        packed = torch.stack([val, idx.float()], -1)            # [B,2]
        allp = mappings.all_gather(packed.unsqueeze(0), 0, g)    # [tp,B,2]
        win = allp[..., 0].argmax(0)                             # lowest rank wins ties, like a flat argmax
        return allp[win, torch.arange(logits.shape[0], device=logits.device), 1].long()

    def forward(self, logits: torch.Tensor, sampling_params: Optional[torch.Tensor] = None,
                rand: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = logits.shape[0]
        if not (self.do_sample or self.dynamic):
            return self.greedy(logits)          # static greedy: no params needed
        if sampling_params is None or not self.dynamic:
            sampling_params = prepare_sampling_params(B, self.cfg.top_k, self.cfg.top_p,
                                                      self.cfg.temperature).to(logits.device)
        top_k = sampling_params[:, 0].to(torch.int32)
        top_p = sampling_params[:, 1]
        temp = sampling_params[:, 2]
        if rand is None:
            rand = self._rand(B, logits.device)
        g = self.tp_group
        if self.vocab_shard:
            K = min(self.global_topk, logits.shape[-1])
            v, i = torch.topk(logits.float(), K, dim=-1)
            i = i + g.rank * logits.shape[-1]
            v = mappings.all_gather(v, -1, g)
            i = mappings.all_gather(i, -1, g)
            tok_local = ops.sample(v, top_k, top_p, temp, rand, self.global_topk)
            return i.gather(1, tok_local.view(B, 1)).view(B)
        return ops.sample(logits, top_k, top_p, temp, rand, self.global_topk)


class DataParallelSampler(Sampler):
    """Sampling DP (reference sampling.py:467-579): instead of every rank reducing the vocab-sharded logits of ALL batch rows,
    an all-to-all hands rank r the full-vocab logits of rows ``r::tp`` — each rank samples ``B/tp`` rows locally (one top-k
    kernel over the whole vocabulary, no staged distributed top-k) and the tokens are all-gathered back."""

    def forward(self, logits: torch.Tensor, sampling_params: Optional[torch.Tensor] = None,
                rand: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = self.tp_group
        B, Vl = logits.shape
        if g.size == 1 or not self.vocab_shard or B % g.size != 0:
            return super().forward(logits, sampling_params, rand)
        n = B // g.size
        # [B, V/tp] -> rows grouped by destination rank -> all-to-all -> [tp(source shard), n, V/tp] -> [n, V]
        send = logits.view(g.size, n, Vl).contiguous()
        recv = mappings.all_to_all(send, 0, 0, g)
        full = recv.permute(1, 0, 2).reshape(n, g.size * Vl)
        sl = slice(g.rank * n, (g.rank + 1) * n)
        sp = None if sampling_params is None else sampling_params[sl]
        rd = None if rand is None else rand[sl]
        vs, self.vocab_shard = self.vocab_shard, False
        try:
            toks = super().forward(full, sp, rd)
        finally:
            self.vocab_shard = vs
        return mappings.all_gather(toks.contiguous(), 0, g)


def create_sampler(neuron_config, tp_group=None, vocab_shard=True):
    if getattr(neuron_config.on_device_sampling_config, "sampling_dp_degree", 1) > 1:
        return DataParallelSampler(neuron_config, tp_group, vocab_shard)
    return Sampler(neuron_config, tp_group, vocab_shard)
