"""Flash decoding: the KV cache of a replicated KV head is *sequence-sharded* across the ranks that replicate it.

reference: modules/flashdecode/utils.py (``calculate_num_cores_per_group``, ``mask_util`` / ``turn_2d_mask_to_4d``, rank-shifted
positions) + the decode path of attention_base.py (all-gather q inside the KV group, local softmax statistics, distributed
log-sum-exp combine).  With ``tp > num_kv_heads`` every KV head lives on ``r = tp / num_kv_heads`` ranks; instead of r full
copies of the sequence each rank keeps the positions ``p % r == j`` (interleaved, so every rank owns the same share of any
prefix), which divides both KV memory and decode attention traffic by r.

Per decode step and layer (one KV group of r ranks, NVLink all-gathers of a few KB):
  1. all-gather the q heads of the group — every rank scores ALL the group's q heads against ITS sequence shard;
  2. the new token's K/V (identical on all r ranks: replicated projection) is written only by the owner ``p % r == j``
     at local slot ``p // r``;
  3. local attention returns un-normalised ``(o, m, l)``;
  4. all-gather ``(o, m, l)``, merge with the usual max/renormalise rule, keep this rank's own q heads."""
from __future__ import annotations

from typing import Tuple

import torch

from ...parallel import mappings


def calculate_num_cores_per_group(num_attention_heads: int, num_key_value_heads: int, tp_degree: int) -> int:
    return tp_degree // num_key_value_heads if tp_degree > num_key_value_heads else 1


def local_slots(positions: torch.Tensor, rank_in_group: int, group_size: int) -> torch.Tensor:
    """Global write positions -> local slot (``-1`` when another rank owns the position; padding stays ``-1``)."""
    own = (positions >= 0) & (positions % group_size == rank_in_group)
    return torch.where(own, positions // group_size, torch.full_like(positions, -1))


def local_horizon(positions: torch.Tensor, rank_in_group: int, group_size: int) -> torch.Tensor:
    """Largest local slot visible to a query at global position P: slots s with ``s*r + j <= P`` (``-1`` = none)."""
    return torch.div(positions - rank_in_group, group_size, rounding_mode="floor")


def partial_attention(q, k, v, horizon, scale) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """q [B,T,Hq,D]; k/v [B,Hkv,S,D] (local shard); horizon [B,T].  -> un-normalised o [B,T,Hq,D] fp32, m, l [B,T,Hq]."""
    B, T, Hq, D = q.shape
    rep = Hq // k.shape[1]
    kk = k.repeat_interleave(rep, 1).float()
    vv = v.repeat_interleave(rep, 1).float()
    s = torch.einsum("bthd,bhsd->bhts", q.float(), kk) * scale
    vis = torch.arange(k.shape[2], device=q.device).view(1, 1, 1, -1) <= horizon.view(B, 1, T, 1)
    s = s.masked_fill(~vis, float("-inf"))
    m = s.amax(-1)
    m_safe = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    e = torch.exp(s - m_safe.unsqueeze(-1))
    e = torch.where(vis, e, torch.zeros_like(e))
    o = torch.einsum("bhts,bhsd->bthd", e, vv)
    return o, m.permute(0, 2, 1), e.sum(-1).permute(0, 2, 1)


def combine(o, m, l, group) -> torch.Tensor:
    """Merge the partial results of the ``group`` ranks.  -> normalised o [B,T,Hq,D] fp32."""
    if group.size == 1:
        return o / l.clamp_min(1e-30).unsqueeze(-1)
    O = mappings.all_gather(o.unsqueeze(0).contiguous(), 0, group)          # [r,B,T,H,D]
    M = mappings.all_gather(m.unsqueeze(0).contiguous(), 0, group)
    L = mappings.all_gather(l.unsqueeze(0).contiguous(), 0, group)
    gm = M.amax(0)
    gm = torch.where(torch.isinf(gm), torch.zeros_like(gm), gm)
    w = torch.exp(torch.where(torch.isinf(M), torch.full_like(M, float("-inf")), M - gm))
    w = torch.where(torch.isinf(M), torch.zeros_like(w), w)
    num = (O * w.unsqueeze(-1)).sum(0)
    den = (L * w).sum(0).clamp_min(1e-30)
    return num / den.unsqueeze(-1)
