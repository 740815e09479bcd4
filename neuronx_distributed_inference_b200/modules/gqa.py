"""GQA head sharding + fused QKV / O projections.

Behaviour of reference modules/attention/gqa.py (sharding strategies :32-100, pad/replicate
:137-243, GroupQueryAttention_QKV :348-953, GroupQueryAttention_O :955-1129) re-expressed as an
*index plan*: for every TP rank the plan lists which source Q heads (or -1 = zero pad) and which
source KV heads it owns.  Sharding a weight, a bias or a per-channel quantisation scale is then
the same row/column gather — no separate pad/replicate code paths for scales.
"""
from __future__ import annotations

import enum
import logging
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..parallel.layers import BaseParallelLinear, _mark
from ..parallel.state import Group, get_tensor_model_parallel_group
from ..parallel import mappings

logger = logging.getLogger("b200infer")


class GQA(enum.Enum):
    CONVERT_TO_MHA = "convert-to-mha"
    REPLICATE_TO_TP_DEGREE = "replicate-to-tp-degree"


def determine_sharding_strategy(tp_degree: int, source_key_value_heads: int,
                                desired_sharding_strategy: Optional[GQA] = None) -> GQA:
    s = desired_sharding_strategy or GQA.REPLICATE_TO_TP_DEGREE
    if s == GQA.REPLICATE_TO_TP_DEGREE and tp_degree % source_key_value_heads != 0:
        if source_key_value_heads % tp_degree == 0:
            return GQA.CONVERT_TO_MHA  # kv heads split evenly: the strategy is moot
        logger.warning("TP degree (%d) and KV heads (%d) are not divisible: using CONVERT_TO_MHA",
                       tp_degree, source_key_value_heads)
        s = GQA.CONVERT_TO_MHA
    return s


def get_number_of_extra_heads(num_heads: int, tp_degree: int) -> int:
    return (-num_heads) % tp_degree


def get_shardable_head_counts(tp_degree: int, num_attention_heads: int, num_key_value_heads: int,
                              sharding_strategy: GQA) -> Tuple[int, int]:
    """Total (padded/replicated) head counts across all ranks."""
    q = num_attention_heads + get_number_of_extra_heads(num_attention_heads, tp_degree)
    kv = num_key_value_heads
    if num_attention_heads == num_key_value_heads:
        kv = q
    elif num_key_value_heads < tp_degree or num_key_value_heads % tp_degree != 0:
        if sharding_strategy == GQA.REPLICATE_TO_TP_DEGREE:
            assert tp_degree % num_key_value_heads == 0
            kv = tp_degree
        else:
            kv = q
    return q, kv


@dataclass
class GQAPlan:
    tp: int
    n_q: int
    n_kv: int
    q_idx: List[List[int]]    # per rank: source q head ids (-1 = pad)
    kv_idx: List[List[int]]   # per rank: source kv head ids

    @property
    def q_per_rank(self):
        return len(self.q_idx[0])

    @property
    def kv_per_rank(self):
        return len(self.kv_idx[0])


def make_gqa_plan(tp: int, n_q: int, n_kv: int, strategy: Optional[GQA] = None) -> GQAPlan:
    strategy = determine_sharding_strategy(tp, n_kv, strategy) if n_q != n_kv else GQA.CONVERT_TO_MHA
    q_tot, kv_tot = get_shardable_head_counts(tp, n_q, n_kv, strategy)
    qpr, kpr = q_tot // tp, kv_tot // tp
    group = n_q // n_kv
    if n_q == n_kv or (kv_tot == q_tot and n_q != n_kv):
        # MHA (or converted to MHA): tail-pad q heads; kv head follows its q head
        flat_q = list(range(n_q)) + [-1] * (q_tot - n_q)
        q_idx = [flat_q[r * qpr:(r + 1) * qpr] for r in range(tp)]
        kv_idx = [[(h // group if h >= 0 else 0) for h in qs] for qs in q_idx]
        return GQAPlan(tp, n_q, n_kv, q_idx, kv_idx)
    if kv_tot == n_kv:
        # enough KV heads: contiguous split of KV heads and their q groups
        assert n_q % tp == 0, "num_attention_heads must divide by tp when kv heads are not replicated"
        kv_idx = [list(range(r * kpr, (r + 1) * kpr)) for r in range(tp)]
        q_idx = [[k * group + j for k in ks for j in range(group)] for ks in kv_idx]
        return GQAPlan(tp, n_q, n_kv, q_idx, kv_idx)
    # REPLICATE_TO_TP_DEGREE: one kv head per rank, q heads of a group interleave-padded
    rep = tp // n_kv
    tgt_group = q_tot // n_kv
    q_idx, kv_idx = [], []
    for r in range(tp):
        g, j = divmod(r, rep)
        heads = []
        for t in range(j * qpr, (j + 1) * qpr):
            heads.append(g * group + t if t < group else -1)
        q_idx.append(heads)
        kv_idx.append([g])
    assert tgt_group == rep * qpr
    return GQAPlan(tp, n_q, n_kv, q_idx, kv_idx)


def _gather_heads(t: torch.Tensor, idx: List[int], head_dim: int, dim: int) -> torch.Tensor:
    """Pick head blocks ``idx`` (size head_dim each, -1 -> zeros) along ``dim``."""
    shape = list(t.shape)
    n_src = shape[dim] // head_dim
    v = t.reshape(shape[:dim] + [n_src, head_dim] + shape[dim + 1:])
    sel = torch.tensor([max(i, 0) for i in idx], dtype=torch.long)
    cast = v.dtype in (torch.float8_e4m3fn, torch.float8_e5m2)
    out = (v.view(torch.uint8) if cast else v).index_select(dim, sel).clone()
    pad = [k for k, i in enumerate(idx) if i < 0]
    if pad:
        out.index_fill_(dim, torch.tensor(pad, dtype=torch.long), 0)
    if cast:
        out = out.view(t.dtype)
    shape[dim] = len(idx) * head_dim
    return out.reshape(shape)


def head_pad_index(src_dim: int, dst_dim: int, split: bool) -> torch.Tensor:
    """Where the ``src_dim`` real channels of a head live inside its zero-padded ``dst_dim`` storage.  ``split`` (rotate-half RoPE):
    the two halves keep their pairing distance — first half at [0, src/2), second half at [dst/2, dst/2 + src/2) — so a kernel that
    rotates channel i with channel i + dst/2 rotates exactly the original pairs; otherwise (interleaved RoPE / no RoPE): a prefix."""
    if not split:
        return torch.arange(src_dim)
    h = src_dim // 2
    return torch.cat([torch.arange(h), dst_dim // 2 + torch.arange(h)])


def _pad_heads(t: torch.Tensor, src_dim: int, dst_dim: int, dim: int, split: bool) -> torch.Tensor:
    """[..., heads * src_dim, ...] -> [..., heads * dst_dim, ...] along ``dim``: each head zero-padded (see head_pad_index)."""
    if src_dim == dst_dim:
        return t
    shape = list(t.shape)
    heads = shape[dim] // src_dim
    cast = t.dtype in (torch.float8_e4m3fn, torch.float8_e5m2)
    v = (t.view(torch.uint8) if cast else t).reshape(shape[:dim] + [heads, src_dim] + shape[dim + 1:])
    out = v.new_zeros(shape[:dim] + [heads, dst_dim] + shape[dim + 1:])
    out.index_copy_(dim + 1, head_pad_index(src_dim, dst_dim, split).to(v.device), v)
    shape[dim] = heads * dst_dim
    out = out.reshape(shape)
    return out.view(t.dtype) if cast else out


class GroupQueryAttention_QKV(BaseParallelLinear):
    """Fused ``Wqkv`` column-parallel projection: local rows = [q heads | k heads | v heads] of this
    rank.  State-dict key: ``Wqkv.weight`` holding the *unsharded* concat [q; k; v]."""

    def __init__(self, hidden_size: int, head_dim: int, num_attention_heads: int, num_key_value_heads: int,
                 tp_group: Optional[Group] = None, dtype=torch.float32, bias: bool = False,
                 desired_sharding_strategy: Optional[GQA] = None, device=None,
                 sequence_parallel_enabled: bool = False, sequence_dimension: int = 1,
                 src_head_dim: Optional[int] = None, pad_split: bool = True):
        """``src_head_dim``: head size of the CHECKPOINT when the layer stores its heads zero-padded to ``head_dim`` (odd head sizes
        such as 80 / 96 / 100 run on the 64 / 128-wide attention kernels; AttentionBase decides)."""
        super().__init__()
        self.tensor_parallel_group = tp_group or get_tensor_model_parallel_group()
        tp = self.tensor_parallel_group.size
        self.hidden_size, self.head_dim = hidden_size, head_dim
        self.src_head_dim, self.pad_split = src_head_dim or head_dim, pad_split
        self.plan = make_gqa_plan(tp, num_attention_heads, num_key_value_heads, desired_sharding_strategy)
        self.n_q, self.n_kv = self.plan.q_per_rank, self.plan.kv_per_rank
        self.sequence_parallel_enabled = sequence_parallel_enabled
        self.sequence_dimension = sequence_dimension
        out_local = (self.n_q + 2 * self.n_kv) * head_dim
        self.weight = _mark(nn.Parameter(torch.empty(out_local, hidden_size, dtype=dtype, device=device),
                                         requires_grad=False), 0, self.tensor_parallel_group)
        self.weight.shard_fn = self._shard
        if bias:
            self.bias = _mark(nn.Parameter(torch.zeros(out_local, dtype=dtype, device=device), requires_grad=False),
                              0, self.tensor_parallel_group)
            self.bias.shard_fn = self._shard
        else:
            self.register_parameter("bias", None)

    def _shard(self, full: torch.Tensor, rank: int) -> torch.Tensor:
        """full: [(n_q + 2 n_kv) * D, ...] unsharded (works for weight, bias, per-channel scale)."""
        p, D, Dp, sp = self.plan, self.src_head_dim, self.head_dim, self.pad_split
        if full.shape[0] == 1:  # per-tensor scale
            return full.clone()
        q, k, v = full.split([p.n_q * D, p.n_kv * D, p.n_kv * D], 0)
        return torch.cat([_pad_heads(_gather_heads(q, p.q_idx[rank], D, 0), D, Dp, 0, sp),
                          _pad_heads(_gather_heads(k, p.kv_idx[rank], D, 0), D, Dp, 0, sp),
                          _pad_heads(_gather_heads(v, p.kv_idx[rank], D, 0), D, Dp, 0, sp)], 0)

    def zero_head_padding(self):
        """Random-init helper: the padded channels of every head must be zero (they are, after a checkpoint load)."""
        if self.src_head_dim == self.head_dim:
            return
        keep = torch.zeros(self.head_dim, dtype=torch.bool)
        keep[head_pad_index(self.src_head_dim, self.head_dim, self.pad_split)] = True
        rows = keep.repeat(self.n_q + 2 * self.n_kv).to(self.weight.device)
        with torch.no_grad():
            for t in (self.weight, self.bias):
                if t is not None:
                    z = t.view(torch.uint8) if t.dtype in (torch.float8_e4m3fn, torch.float8_e5m2) else t
                    z[~rows] = 0

    def preshard_hook(self, model_state_dict: dict, prefix: str) -> bool:
        """Rewrite the fused ``<prefix>.weight / .bias / .scale`` entries of an UNSHARDED state dict to the replicated / padded head
        layout of this TP degree: rank ``r``'s shard becomes the ``r``-th equal row block, so a generic dim-0 splitter (offline sharded
        checkpoints, ``save_sharded_checkpoint``) produces exactly what ``shard_fn`` would (reference gqa.py preshard_hook: KV-head
        replication, Q-head padding, applied to weights AND per-channel quantisation scales)."""
        tp = self.tensor_parallel_group.size
        base = prefix[: -len(".weight")] if prefix.endswith(".weight") else prefix.rstrip(".")
        done = False
        for suffix in ("weight", "bias", "scale"):
            k = f"{base}.{suffix}"
            t = model_state_dict.get(k)
            if t is None or t.shape[0] == 1:
                continue
            model_state_dict[k] = torch.cat([self._shard(t, r) for r in range(tp)], 0)
            done = True
        return done

    def forward(self, x, norm_weight=None, norm_eps=1e-6, norm_offset=0.0):
        if self.sequence_parallel_enabled:
            x = mappings.all_gather(x, self.sequence_dimension, self.tensor_parallel_group)
        return ops.linear(x, self.weight, self.bias, norm_weight=norm_weight, norm_eps=norm_eps,
                          norm_offset=norm_offset, scale=getattr(self, "scale", None))


class GroupQueryAttention_O(BaseParallelLinear):
    """Row-parallel output projection whose input columns follow the q-head plan."""

    def __init__(self, hidden_size: int, head_dim: int, num_attention_heads: int, num_key_value_heads: int,
                 tp_group: Optional[Group] = None, dtype=torch.float32, bias: bool = False,
                 desired_sharding_strategy: Optional[GQA] = None, device=None,
                 sequence_parallel_enabled: bool = False, sequence_dimension: int = 1,
                 reduce_dtype=None, out_size: Optional[int] = None, src_head_dim: Optional[int] = None, pad_split: bool = True):
        super().__init__()
        self.tensor_parallel_group = tp_group or get_tensor_model_parallel_group()
        tp = self.tensor_parallel_group.size
        self.plan = make_gqa_plan(tp, num_attention_heads, num_key_value_heads, desired_sharding_strategy)
        self.head_dim = head_dim
        self.src_head_dim, self.pad_split = src_head_dim or head_dim, pad_split
        self.sequence_parallel_enabled = sequence_parallel_enabled
        self.sequence_dimension = sequence_dimension
        self.reduce_dtype = reduce_dtype
        out_size = out_size or hidden_size
        self.weight = _mark(nn.Parameter(torch.empty(out_size, self.plan.q_per_rank * head_dim, dtype=dtype,
                                                     device=device), requires_grad=False),
                            1, self.tensor_parallel_group)
        self.weight.shard_fn = self._shard
        if bias:
            self.bias = _mark(nn.Parameter(torch.zeros(out_size, dtype=dtype, device=device), requires_grad=False),
                              None, self.tensor_parallel_group)
        else:
            self.register_parameter("bias", None)

    def _shard(self, full: torch.Tensor, rank: int) -> torch.Tensor:
        if full.dim() == 1 or full.shape[-1] == 1:  # per-out-channel scale / per-tensor: replicated
            return full.clone()
        return _pad_heads(_gather_heads(full, self.plan.q_idx[rank], self.src_head_dim, 1), self.src_head_dim, self.head_dim, 1,
                          self.pad_split)

    def zero_head_padding(self):
        if self.src_head_dim == self.head_dim:
            return
        keep = torch.zeros(self.head_dim, dtype=torch.bool)
        keep[head_pad_index(self.src_head_dim, self.head_dim, self.pad_split)] = True
        cols = keep.repeat(self.plan.q_per_rank).to(self.weight.device)
        with torch.no_grad():
            z = self.weight.view(torch.uint8) if self.weight.dtype in (torch.float8_e4m3fn, torch.float8_e5m2) else self.weight
            z[:, ~cols] = 0

    def forward(self, x, residual=None):
        g = self.tensor_parallel_group
        if g.size == 1:
            return ops.linear(x, self.weight, self.bias, residual=residual, scale=getattr(self, "scale", None))
        if self.sequence_parallel_enabled:
            y = mappings.reduce_scatter(ops.linear(x, self.weight, None, scale=getattr(self, "scale", None)),
                                        self.sequence_dimension, g)
            if self.bias is not None:
                y = y + self.bias
            return y if residual is None else y + residual
        return ops.linear_allreduce(x, self.weight, self.bias, g, residual=residual, reduce_dtype=self.reduce_dtype,
                                    scale=getattr(self, "scale", None))
