"""Attention helpers under the reference's names (modules/attention/utils.py): head layout moves, KV repetition, rotary
application, the fp32 "manual" softmax over prior + active scores, distributed softmax statistics, mask builders."""
from __future__ import annotations

import torch

from ... import ops
from ...ops import reference as ref
from ..rope import RotaryEmbedding  # noqa: F401


def move_heads_front(t: torch.Tensor, bsz: int, seq_len: int, num_head: int, head_dim: int, layernorm=None) -> torch.Tensor:
    """[B, S, H*D] -> [B, H, S, D] (optionally per-head layernorm first)."""
    t = t.view(bsz, seq_len, num_head, head_dim)
    if layernorm is not None:
        t = layernorm(t)
    return t.permute(0, 2, 1, 3).contiguous()


def repeat_kv(hidden_states: torch.Tensor, n_rep: int) -> torch.Tensor:
    """[B, Hkv, S, D] -> [B, Hkv*n_rep, S, D]."""
    return hidden_states if n_rep == 1 else hidden_states.repeat_interleave(n_rep, dim=1)


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin, position_ids=None, unsqueeze_dim: int = 1):
    """HF convention: q/k [B, H, S, D], cos/sin [B, S, D] (full width)."""
    cos, sin = cos.unsqueeze(unsqueeze_dim), sin.unsqueeze(unsqueeze_dim)
    return (q * cos + rotate_half(q) * sin), (k * cos + rotate_half(k) * sin)


def manual_softmax(prior_scores, active_scores, is_speculation: bool = False):
    """Joint softmax over the cached ("prior") and the new ("active") keys, fp32 (reference attention/utils.py)."""
    mx = torch.maximum(prior_scores.amax(-1, keepdim=True), active_scores.amax(-1, keepdim=True))
    ep, ea = torch.exp(prior_scores.float() - mx), torch.exp(active_scores.float() - mx)
    den = ep.sum(-1, keepdim=True) + ea.sum(-1, keepdim=True)
    return ep / den, ea / den


def distributed_softmax(prior_scores, active_scores, group=None):
    """Flash-decoding flavour: max / sum statistics are all-reduced across the KV-shard group."""
    from ...parallel import mappings
    from ...parallel.state import get_kv_shared_group
    g = group or get_kv_shared_group()
    mx = torch.maximum(prior_scores.amax(-1, keepdim=True), active_scores.amax(-1, keepdim=True)).float()
    if g.size > 1:
        mx = mappings.all_gather(mx.unsqueeze(0), 0, g).amax(0)
    ep, ea = torch.exp(prior_scores.float() - mx), torch.exp(active_scores.float() - mx)
    den = ep.sum(-1, keepdim=True) + ea.sum(-1, keepdim=True)
    if g.size > 1:
        den = mappings.all_reduce(den, g)
    return ep / den, ea / den


def create_block_diagonal_attn_mask(query_lens, key_lens, max_query_len: int, max_key_len: int) -> torch.Tensor:
    """Chunked-prefill mask: sequence i's queries see only sequence i's keys, causally aligned at the end."""
    m = torch.zeros(max_query_len, max_key_len, dtype=torch.bool)
    qo = ko = 0
    for ql, kl in zip(query_lens.tolist(), key_lens.tolist()):
        for i in range(ql):
            m[qo + i, ko: ko + kl - ql + i + 1] = True
        qo, ko = qo + ql, ko + kl
    return m


build_mask = ref.build_mask
attention_with_mask = ops.attention_with_mask


def get_last_kv_window(window_size: int, position_ids: torch.Tensor, k: torch.Tensor, v: torch.Tensor, windowed_context_encoding: bool = False):
    """Last ``window_size`` keys/values of every row (k/v [B,H,S,D]; rows right padded, last real token = max position)."""
    B, H, S, D = k.shape
    last = position_ids.long().amax(-1)                                   # [B]
    start = (last + 1 - window_size).clamp_min(0)
    idx = (start.view(B, 1) + torch.arange(window_size, device=k.device).view(1, -1)).clamp_max(S - 1)
    gi = idx.view(B, 1, window_size, 1).expand(B, H, window_size, D)
    return k.gather(2, gi), v.gather(2, gi)


def get_last_kv_chunk(chunk_size: int, position_ids: torch.Tensor, k: torch.Tensor, v: torch.Tensor):
    """Keys/values of the chunk that contains each row's last token (Llama-4 chunked attention)."""
    B, H, S, D = k.shape
    last = position_ids.long().amax(-1)
    start = (last // chunk_size) * chunk_size
    idx = (start.view(B, 1) + torch.arange(chunk_size, device=k.device).view(1, -1)).clamp_max(S - 1)
    gi = idx.view(B, 1, chunk_size, 1).expand(B, H, chunk_size, D)
    return k.gather(2, gi), v.gather(2, gi)


def stride_tensor(t: torch.Tensor, dim: int, stride: int) -> torch.Tensor:
    """Strided context parallelism: reorder so that rank r's contiguous shard holds positions r, r+stride, ... (load balance of
    causal attention across CP ranks)."""
    n = t.shape[dim]
    order = torch.arange(n, device=t.device).view(n // stride, stride).t().reshape(-1)
    return t.index_select(dim, order)


def order_strided_tensor(t: torch.Tensor, dim: int, stride: int) -> torch.Tensor:
    """Inverse of :func:`stride_tensor`."""
    n = t.shape[dim]
    order = torch.arange(n, device=t.device).view(stride, n // stride).t().reshape(-1)
    return t.index_select(dim, order)


def validate_tp_prefill_to_dp_decode(num_kv_heads: int, world_size: int, dp_degree: int) -> None:
    """Prefill in full TP, decode attention in TP x DP (reference :603-622): legal when a rank holds the SAME number of KV heads in
    both modes, i.e. the DP groups are exactly the ranks that replicate a KV head — otherwise prefill would have to all-gather KV
    heads into the decode layout.  Same constraint as ``attention_dp_degree == tp / num_kv_heads``."""
    tp_decode = world_size // dp_degree
    dp_heads = max(num_kv_heads // tp_decode, 1)
    tp_heads = max(num_kv_heads // world_size, 1)
    if dp_heads != tp_heads:
        raise ValueError(f"attention DP degree {dp_degree} is not supported for {num_kv_heads} KV heads on {world_size} ranks: a rank "
                         f"would own {dp_heads} KV heads in decode but {tp_heads} in prefill")
