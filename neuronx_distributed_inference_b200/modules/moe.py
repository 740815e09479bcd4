"""Mixture-of-experts block: router + routed experts (+ shared experts), tensor- and expert-parallel.

Rebuild of the external ``neuronx_distributed.modules.moe`` surface used by the reference (SURVEY §2.10:
``RouterTopK``, ``ExpertMLPs(V2)``, ``SharedExperts``, ``MoE``; glue in modules/moe.py:6-45, moe_v2.py:23-161).

Parallel layout on B200 (one NVSwitch domain, activations replicated across the TP group as in every TP decoder):
  * experts are split over ``moe_ep`` groups (contiguous ownership), each expert's intermediate dim over ``moe_tp``;
    ``moe_ep * moe_tp == tp_degree``;
  * every rank evaluates only its (expert, I-slice) shard on the replicated tokens and contributes a partial
    ``[N, H]``; the combine is one all-reduce over the TP group — on the decode path it is the fused
    GEMV->all-reduce epilogue, at prefill sizes the in-switch all-reduce.  This is the reference's
    ``ep_dispatch_cc_option="AR_AG"`` family; neither the reference nor this engine shards the TOKENS across expert ranks, so
    there is no token all-to-all.  Routed experts run through ``ops.moe_experts`` (device-side permutation + grouped tcgen05 GEMMs).
Weights are stored K-major: ``gate_up_proj [E_local, 2*I_local, H]`` ([gate; up] rows) and ``down_proj
[E_local, H, I_local]`` (the reference keeps ``[E,H,2I]`` / ``[E,I,H]``; conversion happens at checkpoint load).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn as nn

from .. import ops
from ..parallel import mappings
from ..parallel.state import (Group, get_expert_model_parallel_group, get_moe_tp_group,
                              get_tensor_model_parallel_group)


class RouterTopK(nn.Module):
    """fp32 router (reference uses fp32 softmax for Qwen3-MoE / sigmoid for Llama-4)."""

    def __init__(self, num_experts: int, top_k: int, hidden_size: int, dtype=torch.float32, act_fn: str = "softmax",
                 bias: bool = False, normalize_top_k_affinities: bool = True, apply_act_fn_over_topk: bool = False,
                 device=None):
        super().__init__()
        self.num_experts, self.top_k, self.act_fn = num_experts, top_k, act_fn
        self.normalize = normalize_top_k_affinities
        self.act_over_topk = apply_act_fn_over_topk
        self.linear_router = nn.Linear(hidden_size, num_experts, bias=bias, dtype=torch.float32, device=device)
        for p in self.linear_router.parameters():
            p.requires_grad_(False)
        self.compute_dtype = torch.float32      # router_config["dtype"] (reference RouterConfig): precision of the router GEMM

    def forward(self, x: torch.Tensor):
        if self.compute_dtype != torch.float32:
            cd = self.compute_dtype
            logits = nn.functional.linear(x.to(cd), self.linear_router.weight.to(cd),
                                          None if self.linear_router.bias is None else self.linear_router.bias.to(cd)).float()
            w, idx = ops.moe_route(logits, self.top_k, self.act_fn, self.normalize, self.act_over_topk)
            return logits, w, idx
        logits = nn.functional.linear(x.float(), self.linear_router.weight, self.linear_router.bias)
        w, idx = ops.moe_route(logits, self.top_k, self.act_fn, self.normalize, self.act_over_topk)
        return logits, w, idx


class ExpertMLPs(nn.Module):
    def __init__(self, num_experts: int, hidden_size: int, intermediate_size: int, hidden_act: str = "silu",
                 dtype=torch.float32, bias: bool = False, device=None, ep_group: Optional[Group] = None,
                 moe_tp_group: Optional[Group] = None, act_fn: Optional[Callable] = None, gated: bool = True):
        """``gated=False``: plain two-matrix experts ``down(act(up(x)))`` (Nemotron-H / Nemotron-3: squared ReLU); ``gate_up_proj`` is then
        just the up projection ``[E, I, H]``."""
        super().__init__()
        self.gated = gated
        self.ep_group = ep_group or get_expert_model_parallel_group()
        self.tp_group = moe_tp_group or get_moe_tp_group()
        ep, tp = self.ep_group.size, self.tp_group.size
        assert num_experts % ep == 0 and intermediate_size % tp == 0
        self.num_experts, self.E_local = num_experts, num_experts // ep
        self.I_local = intermediate_size // tp
        self.expert_offset = self.ep_group.rank * self.E_local
        if gated:
            self.act = {"silu": "silu_mul", "swish": "silu_mul", "gelu": "gelu_mul", "gelu_pytorch_tanh": "gelu_tanh_mul"}[hidden_act]
        else:
            self.act = {"swish": "silu", "gelu_pytorch_tanh": "gelu_tanh"}.get(hidden_act, hidden_act)
        self.act_fn = act_fn
        nproj = 2 if gated else 1
        self.gate_up_proj = nn.Parameter(torch.empty(self.E_local, nproj * self.I_local, hidden_size, dtype=dtype, device=device),
                                         requires_grad=False)
        self.down_proj = nn.Parameter(torch.empty(self.E_local, hidden_size, self.I_local, dtype=dtype, device=device),
                                      requires_grad=False)
        E0, El, tpr, tps = self.expert_offset, self.E_local, self.tp_group.rank, tp

        def shard_gu(full, rank):
            e = full[E0:E0 + El]
            return torch.cat([h.chunk(tps, 1)[tpr] for h in e.chunk(nproj, 1)], 1).contiguous()

        def shard_dn(full, rank):
            return full[E0:E0 + El].chunk(tps, 2)[tpr].contiguous()
        self.gate_up_proj.shard_fn = shard_gu
        self.down_proj.shard_fn = shard_dn
        for p in (self.gate_up_proj, self.down_proj):
            p.tp_group = get_tensor_model_parallel_group()
            p.partition_dim = 0
        if bias:
            self.gate_up_bias = nn.Parameter(torch.zeros(self.E_local, nproj * self.I_local, dtype=dtype, device=device),
                                             requires_grad=False)
            self.down_bias = nn.Parameter(torch.zeros(self.E_local, hidden_size, dtype=dtype, device=device), requires_grad=False)
            self.gate_up_bias.shard_fn = lambda full, rank: torch.cat(
                [h.chunk(tps, 1)[tpr] for h in full[E0:E0 + El].chunk(nproj, 1)], 1).contiguous()
            # the down bias must be added once: only the first I-shard carries it
            self.down_bias.shard_fn = lambda full, rank: (full[E0:E0 + El] if tpr == 0 else torch.zeros_like(full[E0:E0 + El]))
            for p in (self.gate_up_bias, self.down_bias):
                p.tp_group = get_tensor_model_parallel_group()
                p.partition_dim = 0
        else:
            self.gate_up_bias = self.down_bias = None
        self.gate_up_scale = self.down_scale = None        # set by quantization.convert (weight-only int8 / fp8 experts)

    def forward(self, x2: torch.Tensor, topk_w: torch.Tensor, topk_i: torch.Tensor, scale_input: bool = False) -> torch.Tensor:
        return ops.moe_experts(x2, self.gate_up_proj, self.down_proj, topk_w, topk_i, self.act, self.expert_offset,
                               self.gate_up_bias, self.down_bias, self.act_fn, scale_input, self.gate_up_scale, self.down_scale)


class MoE(nn.Module):
    """router -> routed experts (+ shared experts) -> combine.  ``forward`` has the GatedMLP calling convention
    (fused input norm, residual) so it drops into :class:`DecoderLayer`."""

    def __init__(self, router: RouterTopK, expert_mlps: ExpertMLPs, shared_experts: Optional[nn.Module] = None,
                 return_router_logits: bool = False, return_expert_index: bool = False,
                 early_affinity_modulation: bool = False):
        super().__init__()
        self.router = router
        self.expert_mlps = expert_mlps
        self.shared_experts = shared_experts
        self.return_router_logits = return_router_logits
        self.return_expert_index = return_expert_index
        self.early_affinity_modulation = early_affinity_modulation   # Llama-4: scale the expert INPUT by the affinity
        self.tp_all = get_tensor_model_parallel_group()
        self.shared_expert_gate = None     # optional [1,H] linear: sigmoid gate on the shared expert (Qwen2-MoE)
        self.last_router_logits = None
        self.last_expert_index = None

    def forward(self, x, norm_weight=None, norm_eps=1e-6, norm_offset=0.0, residual=None, padding_mask=None):
        shape = x.shape
        xn = ops.rmsnorm(x, norm_weight, norm_eps, norm_offset) if norm_weight is not None else x
        x2 = xn.reshape(-1, shape[-1])
        logits, w, idx = self.router(x2)
        if self.return_router_logits:
            self.last_router_logits = logits
        if self.return_expert_index:
            self.last_expert_index = idx
        if self.early_affinity_modulation:
            y = self.expert_mlps(x2, w.to(torch.float32), idx, scale_input=True)   # Llama-4: y = sum_j expert_j(x * w_j)
        else:
            y = self.expert_mlps(x2, w.to(torch.float32), idx)
        if self.shared_experts is not None:
            sh = self.shared_experts(x2, reduce=False)
            if self.shared_expert_gate is not None:    # same scalar on every rank: commutes with the TP reduction below
                sh = sh * torch.sigmoid(torch.nn.functional.linear(x2.float(), self.shared_expert_gate.weight.float())).to(sh.dtype)
            y = y + sh
        if self.tp_all.size > 1:
            y = mappings.all_reduce(y, self.tp_all)
        y = y.view(shape)
        return y if residual is None else y + residual


class SharedExperts(nn.Module):
    """Always-on gated MLP evaluated on every token; output left un-reduced so that MoE does ONE all-reduce."""

    def __init__(self, hidden_size: int, intermediate_size: int, hidden_act: str = "silu", dtype=torch.float32, device=None):
        super().__init__()
        from ..parallel.layers import ColumnParallelLinear, RowParallelLinear
        from .mlp import _GLU_ACT
        self.act = _GLU_ACT[hidden_act]
        self.gate_up_proj = ColumnParallelLinear(hidden_size, 2 * intermediate_size, bias=False, gather_output=False,
                                                 dtype=dtype, device=device, stride=2)
        self.down_proj = RowParallelLinear(intermediate_size, hidden_size, bias=False, dtype=dtype, device=device,
                                           reduce_output=False)

    def forward(self, x2, reduce: bool = False):
        return self.down_proj(self.gate_up_proj(x2, act=self.act))


def initialize_moe_module(config, device=None, hidden_act: Optional[str] = None, shared: bool = False,
                          router_act: str = "softmax", router_bias: bool = False, expert_bias: bool = False,
                          act_fn=None, apply_act_fn_over_topk: bool = False, intermediate_size: Optional[int] = None,
                          early_affinity_modulation: bool = False, normalize: Optional[bool] = None) -> MoE:
    """Factory with the reference's name (modules/moe.py:6-45, moe_v2.py:23-128)."""
    nc = config.neuron_config
    dt = nc.torch_dtype
    E = getattr(config, "num_local_experts", None) or getattr(config, "num_experts", None) or config.n_routed_experts
    k = config.num_experts_per_tok
    I = intermediate_size or getattr(config, "moe_intermediate_size", None) or config.intermediate_size
    if normalize is None:
        normalize = getattr(nc, "normalize_top_k_affinities", True)
    if getattr(nc, "router_config_explicit", False):      # --router-act-fn / --router-dtype
        router_act = nc.router_config.get("act_fn", router_act)
    router = RouterTopK(E, k, config.hidden_size, dt, router_act, router_bias, normalize, apply_act_fn_over_topk, device)
    if getattr(nc, "router_config_explicit", False):
        from ..config import to_torch_dtype
        router.compute_dtype = to_torch_dtype(nc.router_config.get("dtype", "float32"))
    experts = ExpertMLPs(E, config.hidden_size, I, hidden_act or config.hidden_act, dt, expert_bias, device, act_fn=act_fn)
    sh = None
    if shared:
        n_sh = getattr(config, "n_shared_experts", 1) or 1
        sh = SharedExperts(config.hidden_size, I * n_sh if hasattr(config, "n_shared_experts") else config.intermediate_size,
                           hidden_act or config.hidden_act, dt, device)
    return MoE(router, experts, sh, getattr(nc, "return_router_logits", False), getattr(nc, "return_expert_index", False),
               early_affinity_modulation)
