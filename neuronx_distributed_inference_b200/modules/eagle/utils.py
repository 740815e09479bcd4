"""Weight-gathered projections for the EAGLE draft (reference modules/eagle/utils.py:65-205 — ``looped_einsum`` over the
``AwsNeuronCollectiveMatmul`` custom call and ``tiled_all_gather_matmul`` with one AllGather + Dot per 4096-wide K tile — used by
``WeightGatheredColumnParallel.forward_wg``, models/llama/modeling_llama.py:211-258).

Use case: a projection whose WEIGHT is sharded (to save memory) but whose input is replicated and whose FULL output every rank needs
(the draft's ``fc`` feature-fusion layer during long prefills).  ``y = x @ all_gather(W_shard)^T`` is evaluated tile by tile along K so
only one ``[out, tile]`` slab of the gathered weight is alive at a time and tile ``i+1``'s gather overlaps tile ``i``'s GEMM.

On B200 the decode path does not use this (a 2H x H weight is 64 MB for an 8B model and is simply replicated — DESIGN.md §8); the
prefill path can, when ``weight_gather_seq_len_threshold`` is set.  The collective is NCCL all-gather on a side stream; the fused
single-kernel AG->GEMM is a round-2 item."""
from __future__ import annotations

from typing import Optional

import torch

from ...parallel import mappings
from ...parallel.layers import ColumnParallelLinear
from ...parallel.state import Group


def tiled_all_gather_matmul(x: torch.Tensor, w_shard: torch.Tensor, group: Optional[Group], tile: int = 4096,
                            bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``x [..., K] @ all_gather(w_shard [N/tp, K], dim=0)^T -> [..., N]`` in K tiles of ``tile`` columns.
    fp32 accumulation across tiles; the gathered slab of the next tile is requested before the current GEMM is issued."""
    tp = 1 if group is None else group.size
    K = x.shape[-1]
    if tp == 1:
        y = torch.nn.functional.linear(x, w_shard)
        return y if bias is None else y + bias
    tiles = [(k0, min(k0 + tile, K)) for k0 in range(0, K, tile)]
    acc = None
    nxt = mappings.all_gather(w_shard[:, tiles[0][0]:tiles[0][1]].contiguous(), 0, group)
    for i, (k0, k1) in enumerate(tiles):
        cur = nxt
        if i + 1 < len(tiles):
            n0, n1 = tiles[i + 1]
            nxt = mappings.all_gather(w_shard[:, n0:n1].contiguous(), 0, group)
        part = torch.matmul(x[..., k0:k1], cur.t().to(x.dtype)).float()
        acc = part if acc is None else acc + part
    y = acc.to(x.dtype)
    return y if bias is None else y + bias


def looped_einsum(x: torch.Tensor, w_shard: torch.Tensor, group: Optional[Group], loops: int = 1) -> torch.Tensor:
    """The reference's other spelling of the same contraction: the N (output) dimension is processed in ``loops`` slices, each slice
    gathering its rows of the weight from all ranks (``[tp, N/tp/loops, K]``) — bounded peak memory along N instead of K."""
    tp = 1 if group is None else group.size
    if tp == 1:
        return torch.nn.functional.linear(x, w_shard)
    n_local = w_shard.shape[0]
    assert n_local % loops == 0
    step = n_local // loops
    outs = []
    for j in range(loops):
        rows = mappings.all_gather(w_shard[j * step:(j + 1) * step].contiguous().unsqueeze(0), 0, group)      # [tp, step, K]
        outs.append(torch.einsum("...k,rnk->...rn", x, rows.to(x.dtype)))                                       # [..., tp, step]
    y = torch.stack(outs, -2)                                                                                   # [..., tp, loops, step]
    return y.reshape(*x.shape[:-1], tp * n_local)


class WeightGatheredColumnParallel(ColumnParallelLinear):
    """Column-parallel layer with a second forward that returns the FULL output on every rank by gathering the weight instead of the
    activations (cheaper when tokens >> out features / tp, i.e. long prefills)."""

    def forward_wg(self, x: torch.Tensor, tile: int = 4096) -> torch.Tensor:
        return tiled_all_gather_matmul(x, self.weight, self.tensor_parallel_group, tile, self.bias_full() if self.bias is not None else None)

    def bias_full(self):
        g = self.tensor_parallel_group
        return self.bias if g.size == 1 else mappings.all_gather(self.bias, 0, g)
