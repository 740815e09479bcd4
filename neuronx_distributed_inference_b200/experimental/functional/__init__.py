"""Functional building blocks (reference experimental/functional/**: qkv_proj, causal attention, token-generation attention
"megakernels" for standard and block KV layouts, o_proj unreduced, gated MLP (fused / unreduced), norms, token-generation MoE
over all experts, context-parallel split/gather helpers).  Stateless functions over explicit weight tensors; each maps to ONE
of the engine's CUDA kernels on a GPU and to the fp32 oracle on CPU."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ... import ops
from ...parallel import mappings
from ...parallel.state import (Group, get_context_parallel_group, get_context_parallel_tp_group,  # noqa: F401
                               get_tensor_model_parallel_group)


# ---- norms ---------------------------------------------------------------------------------------------------------
def rms_norm(x, weight, eps: float = 1e-6, residual=None):
    return ops.rmsnorm(x, weight, eps, 0.0, residual)


# ---- QKV -----------------------------------------------------------------------------------------------------------
def qkv_proj(x, w_qkv, n_q: int, n_kv: int, head_dim: int, bias=None, norm_weight=None, norm_eps: float = 1e-6):
    """x [B,T,H] @ fused [q;k;v] weight -> (q [B,T,n_q,D], k, v [B,T,n_kv,D]); optional fused input RMSNorm."""
    B, T, _ = x.shape
    qkv = ops.linear(x, w_qkv, bias, norm_weight=norm_weight, norm_eps=norm_eps).view(B, T, n_q + 2 * n_kv, head_dim)
    return qkv.split([n_q, n_kv, n_kv], 2)


qkv_kernel = qkv_proj


# ---- attention -------------------------------------------------------------------------------------------------------
def causal_scaled_dot_product_attention(q, k, v, scale: Optional[float] = None, sliding_window: Optional[int] = None):
    """Prefill: q,k,v [B,T,H,D] (GQA allowed) -> [B,T,Hq,D]."""
    return ops.attention_prefill(q, k, v, scale if scale is not None else q.shape[-1] ** -0.5, True, sliding_window)


scaled_dot_product_attention_kernel = causal_scaled_dot_product_attention


def tokengen_attention_megakernel_standard_kv(qkv, cos, sin, k_cache, v_cache, seq_ids, positions, n_q: int, n_kv: int, head_dim: int,
                                              scale: Optional[float] = None):
    """Decode step on the contiguous cache: RoPE + cache append (one kernel) then split-KV flash decode (one kernel)."""
    q = ops.rope_kv_append(qkv, cos, sin, k_cache, v_cache, seq_ids, positions, n_q, n_kv, head_dim, False, None, None, 1e-6)
    return ops.attention_decode(q, k_cache, v_cache, seq_ids, positions, scale if scale is not None else head_dim ** -0.5)


def tokengen_attention_megakernel_block_kv(q, k_new, v_new, k_cache, v_cache, slot_mapping, block_table, positions,
                                           scale: Optional[float] = None):
    ops.paged_kv_append(k_cache, v_cache, k_new, v_new, slot_mapping)
    return ops.paged_attention_decode(q, k_cache, v_cache, block_table, positions, scale if scale is not None else q.shape[-1] ** -0.5)


def o_proj_kernel_unreduced(attn_out, w_o):
    """Row-parallel output projection WITHOUT the all-reduce (the caller fuses the reduction elsewhere)."""
    return ops.linear(attn_out, w_o)


def o_proj_allreduce(attn_out, w_o, residual=None, group: Optional[Group] = None):
    """GEMV -> all-reduce -> +residual in one kernel on the fused path."""
    return ops.linear_allreduce(attn_out, w_o, None, group or get_tensor_model_parallel_group(), residual=residual)


# ---- MLP -----------------------------------------------------------------------------------------------------------
def gated_mlp_fused(x, w_gate_up, w_down, norm_weight=None, norm_eps: float = 1e-6, residual=None, group: Optional[Group] = None):
    """(RMSNorm ->) [gate;up] GEMM with SwiGLU epilogue -> down GEMM -> all-reduce (+residual)."""
    h = ops.linear(x, w_gate_up, None, norm_weight=norm_weight, norm_eps=norm_eps, act="silu_mul")
    return ops.linear_allreduce(h, w_down, None, group or get_tensor_model_parallel_group(), residual=residual)


def gated_mlp(x, w_gate, w_up, w_down, group: Optional[Group] = None):
    return gated_mlp_fused(x, torch.cat([w_gate, w_up], 0), w_down, group=group)


def gated_mlp_kernel_unreduced(x, w_gate_up, w_down, norm_weight=None, norm_eps: float = 1e-6):
    return ops.linear(ops.linear(x, w_gate_up, None, norm_weight=norm_weight, norm_eps=norm_eps, act="silu_mul"), w_down)


# ---- MoE -----------------------------------------------------------------------------------------------------------
def tokengen_moe_megakernel_forward_all_experts(x, router_w, w_gate_up, w_down, top_k: int, normalize: bool = True,
                                                act: str = "softmax"):
    """Router + expert MLPs for a decode batch; experts [E,2I,H] / [E,H,I]."""
    B, T, H = x.shape
    aff, idx = ops.moe_route(ops.ref.linear(x.reshape(-1, H).float(), router_w.float()), top_k, act, normalize)
    return ops.moe_experts(x.reshape(-1, H), w_gate_up, w_down, aff, idx).view(B, T, H)


def tokengen_moe_megakernel_forward_all_experts_with_shared_experts(x, router_w, w_gate_up, w_down, shared_gate_up, shared_down,
                                                                    top_k: int, normalize: bool = True, act: str = "softmax"):
    y = tokengen_moe_megakernel_forward_all_experts(x, router_w, w_gate_up, w_down, top_k, normalize, act)
    return y + gated_mlp_kernel_unreduced(x, shared_gate_up, shared_down)


# ---- context parallel ----------------------------------------------------------------------------------------------------
def split_input_for_context_parallel(x, dim: int = 1, group: Optional[Group] = None):
    g = group or get_context_parallel_group()
    n = x.shape[dim] // g.size
    return x.narrow(dim, g.rank * n, n)


def gather_kv_context_parallel(k, v, dim: int = 1, group: Optional[Group] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    g = group or get_context_parallel_group()
    return mappings.all_gather(k.contiguous(), dim, g), mappings.all_gather(v.contiguous(), dim, g)


get_context_parallel_cp_group = get_context_parallel_group


def initialize_context_parallel_process_groups(tp_degree: int, cp_degree: int):
    from ...parallel.state import initialize_model_parallel
    return initialize_model_parallel(tensor_model_parallel_size=tp_degree, context_parallel_size=cp_degree)
