"""Functional collectives over a :class:`Group` (inference only, no autograd).

These are the *non-fused* paths (prefill-size tensors, CPU/gloo, bring-up, and the baseline
that the fused GEMM+collective kernels are measured against).  Names follow the call sites of
``neuronx_distributed.parallel_layers.mappings`` in the reference
(modules/attention/attention_base.py:49-55, models/model_base.py:16-20).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .state import Group, get_tensor_model_parallel_group


def _g(group: Optional[Group]) -> Group:
    return group if group is not None else get_tensor_model_parallel_group()


def _heap(g: Group):
    """The group's symmetric heap (parallel/symm_heap.py) when the in-switch (NVLS) collectives are available."""
    return getattr(g, "heap", None)


def all_reduce(x: torch.Tensor, group: Optional[Group] = None, op=None, reduce_dtype=None, residual=None) -> torch.Tensor:
    """sum over the group (+ ``residual``, fused on the NVLS path)."""
    g = _g(group)
    if g.size == 1:
        return x if residual is None else x + residual
    heap = _heap(g)
    if (heap is not None and op in (None, dist.ReduceOp.SUM) and reduce_dtype in (None, x.dtype, torch.float32)
            and heap.usable(x)):
        return heap.all_reduce(x, residual)      # two-shot in-switch reduction, fp32 accumulation
    if residual is not None:
        return all_reduce(x, g, op, reduce_dtype) + residual
    op = op or dist.ReduceOp.SUM
    if reduce_dtype is not None and reduce_dtype != x.dtype:
        y = x.to(reduce_dtype)
        dist.all_reduce(y, op=op, group=g.pg)
        return y.to(x.dtype)
    x = x.contiguous()
    dist.all_reduce(x, op=op, group=g.pg)
    return x


def all_gather(x: torch.Tensor, dim: int, group: Optional[Group] = None, transient: bool = False) -> torch.Tensor:
    """``transient``: the caller consumes the result before the next collective of the group (the NVLS path then returns a view
    of the symmetric staging area instead of a private copy)."""
    g = _g(group)
    if g.size == 1:
        return x
    dim = dim % x.dim()
    heap = _heap(g)
    if heap is not None and x.dim() > 0:
        full = list(x.shape)
        full[dim] *= g.size
        if heap.usable(x, dim, full):
            y = heap.all_gather(x, dim)          # every rank multicasts its slice (multimem.st)
            return y if transient else y.clone()
    x = x.contiguous()
    if x.dim() == 0:
        x = x.reshape(1)
    flat = torch.empty((g.size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, x, group=g.pg)
    out = flat.view((g.size,) + tuple(x.shape))
    if dim == 0:
        return out.reshape((-1,) + tuple(x.shape[1:]))
    return out.movedim(0, dim).reshape(x.shape[:dim] + (g.size * x.shape[dim],) + x.shape[dim + 1:])


def reduce_scatter(x: torch.Tensor, dim: int, group: Optional[Group] = None, op=None, residual=None) -> torch.Tensor:
    g = _g(group)
    if g.size == 1:
        return x if residual is None else x + residual
    dim = dim % x.dim()
    assert x.shape[dim] % g.size == 0
    heap = _heap(g)
    if heap is not None and op in (None, dist.ReduceOp.SUM) and heap.usable(x, dim):
        return heap.reduce_scatter(x, dim, residual)     # rank r pulls the switch-reduced slice r (+ residual)
    if residual is not None:
        return reduce_scatter(x, dim, g, op) + residual
    op = op or dist.ReduceOp.SUM
    xs = x.movedim(dim, 0).contiguous()
    out = torch.empty((xs.shape[0] // g.size,) + tuple(xs.shape[1:]), dtype=x.dtype, device=x.device)
    if x.device.type == "cpu":
        # gloo has no reduce_scatter: all_reduce + slice (host-side test path only)
        dist.all_reduce(xs, op=op, group=g.pg)
        out = xs.chunk(g.size, 0)[g.rank].contiguous()
    else:
        dist.reduce_scatter_tensor(out, xs, op=op, group=g.pg)
    return out.movedim(0, dim)


def all_to_all(x: torch.Tensor, split_dim: int, concat_dim: int, group: Optional[Group] = None) -> torch.Tensor:
    g = _g(group)
    if g.size == 1:
        return x
    ins = [t.contiguous() for t in x.chunk(g.size, split_dim)]
    outs = [torch.empty_like(ins[0]) for _ in range(g.size)]
    if x.device.type == "cpu":
        gathered = [torch.empty_like(x.contiguous()) for _ in range(g.size)]
        dist.all_gather(gathered, x.contiguous(), group=g.pg)
        outs = [t.chunk(g.size, split_dim)[g.rank] for t in gathered]
    else:
        dist.all_to_all(outs, ins, group=g.pg)
    return torch.cat(outs, dim=concat_dim)


def scatter_to_region(x: torch.Tensor, dim: int, group: Optional[Group] = None) -> torch.Tensor:
    """Keep my slice of dim (no communication)."""
    g = _g(group)
    if g.size == 1:
        return x
    return x.chunk(g.size, dim)[g.rank].contiguous()


# ---- reference-compatible names ------------------------------------------------------------
def reduce_from_tensor_model_parallel_region(x, process_group=None, reduce_dtype=None):
    return all_reduce(x, process_group, reduce_dtype=reduce_dtype)


def reduce_scatter_to_sequence_parallel_region(x, dim, process_group=None):
    return reduce_scatter(x, dim, process_group)


def gather_from_sequence_parallel_region(x, dim, process_group=None, tile_cc=None):
    return all_gather(x, dim, process_group)


def gather_from_tensor_model_parallel_region_with_dim(x, gather_dim, process_group=None):
    return all_gather(x, gather_dim, process_group)


def reduce_scatter_to_tensor_model_parallel_region_with_dim(x, partition_dim, process_group=None):
    return reduce_scatter(x, partition_dim, process_group)


def scatter_to_tensor_model_parallel_region(x, dim=-1, process_group=None):
    return scatter_to_region(x, dim, process_group)


_gather_along_dim = all_gather


def _gather_along_first_dim(x, process_group=None):
    return all_gather(x, 0, process_group)


def _reduce_scatter_along_dim(x, dim, op=None, process_group=None):
    return reduce_scatter(x, dim, process_group, op)
