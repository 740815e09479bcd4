"""Tensor-parallel layers: ColumnParallelLinear / RowParallelLinear / ParallelEmbedding.

Inferred API of the external ``neuronx_distributed.parallel_layers`` package that every model
of the reference is wired with (call sites modeling_llama.py:354-387,1112-1133,
gqa.py:411-467,998-1010).  B200 design: each process owns its shard as a plain ``nn.Parameter``
in ``nn.Linear`` layout ``[out, in]``; forward runs

* T <= 16 tokens (decode / speculation): the weight-streaming GEMV kernels of ``ops`` — for the
  row-parallel case the all-reduce (+ residual add) is fused into the GEMV epilogue over
  NVLink peer memory (csrc/gemv_allreduce.cu),
* larger T: the tcgen05 GEMM (csrc/gemm_tcgen05.cu) followed by / preceded by the collective.

Every parameter carries ``partition_dim`` / ``tp_group`` metadata consumed by the checkpoint
sharder (modules/checkpoint.py::shard_state_dict) instead of the reference's trace-time
``shard_checkpoint`` walk.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from . import mappings
from .state import Group, get_tensor_model_parallel_group


def _mark(p: nn.Parameter, partition_dim: Optional[int], group: Group, stride: int = 1):
    p.partition_dim = partition_dim
    p.tp_group = group
    p.partition_stride = stride
    return p


def divide(a: int, b: int) -> int:
    assert a % b == 0, f"{a} is not divisible by {b}"
    return a // b


class BaseParallelLinear(nn.Module):
    quantizable = True


class ColumnParallelLinear(BaseParallelLinear):
    """Y = X W^T with W sharded along the output dim.  ``gather_output`` all-gathers Y."""

    def __init__(self, input_size: int, output_size: int, bias: bool = True, gather_output: bool = True,
                 dtype: torch.dtype = torch.float32, device=None, pad: bool = False,
                 sequence_parallel_enabled: bool = False, sequence_dimension: Optional[int] = None,
                 tensor_model_parallel_group: Optional[Group] = None, skip_bias_add: bool = False,
                 stride: int = 1, **unused):
        super().__init__()
        self.tensor_parallel_group = tensor_model_parallel_group or get_tensor_model_parallel_group()
        tp = self.tensor_parallel_group.size
        self.input_size, self.output_size = input_size, output_size
        self.pad_size = 0
        if pad and output_size % tp != 0:
            self.pad_size = tp - output_size % tp
        self.output_size_per_partition = divide(output_size + self.pad_size, tp)
        self.gather_output = gather_output
        self.skip_bias_add = skip_bias_add
        self.sequence_parallel_enabled = sequence_parallel_enabled
        self.sequence_dimension = 1 if sequence_dimension is None else sequence_dimension
        self.weight = _mark(nn.Parameter(torch.empty(self.output_size_per_partition, input_size, dtype=dtype,
                                                     device=device), requires_grad=False),
                            0, self.tensor_parallel_group, stride)
        if bias:
            self.bias = _mark(nn.Parameter(torch.zeros(self.output_size_per_partition, dtype=dtype, device=device),
                                           requires_grad=False), 0, self.tensor_parallel_group, stride)
        else:
            self.register_parameter("bias", None)

    def forward(self, x: torch.Tensor, norm_weight: Optional[torch.Tensor] = None, norm_eps: float = 1e-6,
                norm_offset: float = 0.0, act: Optional[str] = None) -> torch.Tensor:
        """``norm_weight``: fuse an RMSNorm of x into the GEMM prologue.  ``act='silu_mul'``:
        weight rows are [gate; up] and the output is silu(gate)*up (SwiGLU epilogue)."""
        if self.sequence_parallel_enabled:
            # consumed by the GEMM right below: the NVLS path hands back a view of the symmetric staging area (no copy)
            x = mappings.all_gather(x, self.sequence_dimension, self.tensor_parallel_group, transient=True)
        bias = None if self.skip_bias_add else self.bias
        y = ops.linear(x, self.weight, bias, norm_weight=norm_weight, norm_eps=norm_eps,
                       norm_offset=norm_offset, act=act, scale=getattr(self, "scale", None))
        if self.gather_output:
            y = mappings.all_gather(y, -1, self.tensor_parallel_group)
            if self.pad_size:
                y = y[..., : self.output_size]
        return (y, self.bias) if self.skip_bias_add else y


class RowParallelLinear(BaseParallelLinear):
    """Y = sum_ranks X_r W_r^T with W sharded along the input dim; output all-reduced
    (or reduce-scattered along the sequence when sequence parallel)."""

    def __init__(self, input_size: int, output_size: int, bias: bool = True, input_is_parallel: bool = True,
                 dtype: torch.dtype = torch.float32, device=None, pad: bool = False,
                 sequence_parallel_enabled: bool = False, sequence_dimension: Optional[int] = None,
                 tensor_model_parallel_group: Optional[Group] = None, reduce_dtype: Optional[torch.dtype] = None,
                 reduce_output: bool = True, stride: int = 1, **unused):
        super().__init__()
        self.tensor_parallel_group = tensor_model_parallel_group or get_tensor_model_parallel_group()
        tp = self.tensor_parallel_group.size
        self.input_size, self.output_size = input_size, output_size
        self.pad_size = 0
        if pad and input_size % tp != 0:
            self.pad_size = tp - input_size % tp
        self.input_size_per_partition = divide(input_size + self.pad_size, tp)
        self.input_is_parallel = input_is_parallel
        self.sequence_parallel_enabled = sequence_parallel_enabled
        self.sequence_dimension = 1 if sequence_dimension is None else sequence_dimension
        self.reduce_dtype = reduce_dtype
        self.reduce_output = reduce_output
        self.weight = _mark(nn.Parameter(torch.empty(output_size, self.input_size_per_partition, dtype=dtype,
                                                     device=device), requires_grad=False),
                            1, self.tensor_parallel_group, stride)
        if bias:
            # bias is replicated and added once, after the reduction
            self.bias = _mark(nn.Parameter(torch.zeros(output_size, dtype=dtype, device=device),
                                           requires_grad=False), None, self.tensor_parallel_group)
        else:
            self.register_parameter("bias", None)

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``residual`` (same shape as the output) is added after the reduction; on the fused
        decode path this is the GEMV -> all-reduce -> +residual kernel."""
        g = self.tensor_parallel_group
        if not self.input_is_parallel:
            x = mappings.scatter_to_region(x, -1, g)
        if g.size == 1 or not self.reduce_output:
            return ops.linear(x, self.weight, self.bias if g.size == 1 or g.rank == 0 else None, residual=residual,
                              scale=getattr(self, "scale", None))
        if self.sequence_parallel_enabled:
            # the replicated bias joins the partial sums on rank 0 only; the GEMM writes its partials straight into the symmetric
            # staging area when the group has one, and the reduce-scatter (+ residual) is one in-switch kernel
            b0 = self.bias if g.rank == 0 else None
            heap = getattr(g, "heap", None)
            sd = self.sequence_dimension % x.dim()
            if (heap is not None and getattr(self, "scale", None) is None and sd == x.dim() - 2
                    and heap.gemm_rs_usable(x, self.weight, x.shape[sd])):
                # ONE kernel: tcgen05 GEMM -> per-tile hand-off over NVLink -> in-switch reduce (multimem.ld_reduce) of the rows this
                # rank owns -> + residual (csrc/gemm_tcgen05.cu, fused reduce-scatter epilogue)
                ops.stats["gemm_reduce_scatter"] += 1
                y = heap.gemm_reduce_scatter(x.reshape(-1, x.shape[-1]), self.weight, b0, x.shape[sd], residual)
                return y.view(*x.shape[:sd], x.shape[sd] // g.size, self.weight.shape[0])
            y = ops.linear(x, self.weight, b0, scale=getattr(self, "scale", None), out=ops.staging_for(g, x, self.weight))
            return mappings.reduce_scatter(y, self.sequence_dimension, g, residual=residual)
        return ops.linear_allreduce(x, self.weight, self.bias, g, residual=residual,
                                    reduce_dtype=self.reduce_dtype, scale=getattr(self, "scale", None))


class ParallelEmbedding(nn.Module):
    """Embedding table sharded along the hidden dim (``shard_across_embedding=True``: local
    lookup + all-gather) or along the vocab (mask out-of-range ids + all-reduce).
    reference call sites: modeling_llama.py:1112-1123, model_base.py:1500-1514."""

    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None,
                 dtype: torch.dtype = torch.float32, device=None, shard_across_embedding: bool = True,
                 pad: bool = False, tensor_model_parallel_group: Optional[Group] = None,
                 sequence_parallel_enabled: bool = False, sequence_dimension: Optional[int] = None, **unused):
        super().__init__()
        self.tensor_parallel_group = tensor_model_parallel_group or get_tensor_model_parallel_group()
        tp = self.tensor_parallel_group.size
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.padding_idx = padding_idx
        self.shard_across_embedding = shard_across_embedding
        self.sequence_parallel_enabled = sequence_parallel_enabled
        self.sequence_dimension = 1 if sequence_dimension is None else sequence_dimension
        self.pad_size = 0
        # B200: a GPU has 180 GB — sharding a ~1 GB table buys nothing and costs a collective (all-gather / all-reduce) inside
        # every decode step.  On CUDA the table is REPLICATED (partition_dim None: the checkpoint sharder copies it whole) and
        # the lookup is local; NXDI_B200_SHARD_EMBEDDING=1 restores the reference layouts (also what the CPU / gloo path tests).
        self.replicated = (tp > 1 and device is not None and torch.device(device).type == "cuda"
                           and os.environ.get("NXDI_B200_SHARD_EMBEDDING", "0") != "1")
        if self.replicated:
            shape = (num_embeddings, embedding_dim)
            pdim = None
        elif shard_across_embedding:
            shape = (num_embeddings, divide(embedding_dim, tp))
            pdim = 1
        else:
            if pad and num_embeddings % tp != 0:
                self.pad_size = tp - num_embeddings % tp
            per = divide(num_embeddings + self.pad_size, tp)
            shape = (per, embedding_dim)
            pdim = 0
            self.vocab_start = self.tensor_parallel_group.rank * per
            self.vocab_end = self.vocab_start + per
        self.weight = _mark(nn.Parameter(torch.empty(*shape, dtype=dtype, device=device), requires_grad=False),
                            pdim, self.tensor_parallel_group)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        g = self.tensor_parallel_group
        if g.size == 1 or self.replicated:
            return nn.functional.embedding(ids, self.weight)
        if self.shard_across_embedding:
            y = nn.functional.embedding(ids, self.weight)
            return mappings.all_gather(y, -1, g)
        mask = (ids < self.vocab_start) | (ids >= self.vocab_end)
        local = (ids - self.vocab_start).masked_fill(mask, 0)
        y = nn.functional.embedding(local, self.weight)
        y = y.masked_fill(mask.unsqueeze(-1), 0)
        if self.sequence_parallel_enabled:
            return mappings.reduce_scatter(y, self.sequence_dimension, g)
        return mappings.all_reduce(y, g)
