"""MoE v2 surface (reference modules/moe_v2.py:23-161): the config objects the reference hands to the external ``ExpertMLPsV2`` —
``RoutedExpertsMLPOpsConfig`` (expert shapes, GLU flavour, activation scaling / bias, gate / up clamps, affinity handling),
``BlockwiseMatmulConfig`` (prefill token-block mapping) and ``MoEFusedTKGConfig`` (decode mega-kernel switches) — and the factory
``initialize_moe_module`` / ``initialize_moe_process_group``.

B200 mapping of the knobs:
* GLU flavour + ``hidden_act_scaling_factor`` / ``hidden_act_bias`` + clamps -> one elementwise epilogue ``act_fn([gate | up])`` applied
  between the two expert GEMMs (GPT-OSS: ``(up + 1) * gate * sigmoid(1.702 * gate)`` with gate <= 7, |up| <= 7).
* ``BlockwiseMatmulConfig.block_size`` -> token tile of the grouped prefill GEMM; the NKI-only switches are accepted and ignored.
* ``MoEFusedTKGConfig.moe_fused_kernel_enabled`` / ``expert_mlp_kernel_enabled`` -> whether decode may take the streaming
  ``moe_decode`` CUDA kernel (csrc/moe_decode.cu); off = batched-GEMM reference path.
* TP x EP groups come from ``parallel.state`` (``initialize_moe_process_group`` builds them from the hybrid sharding config)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch

from .moe import MoE, ExpertMLPs, RouterTopK, SharedExperts, initialize_moe_module as _initialize_v1  # noqa: F401


@dataclass
class RoutedExpertsMLPOpsConfig:
    num_experts: int
    hidden_size: int
    intermediate_size: int
    top_k: int
    hidden_act: str = "silu"
    bias: bool = False
    glu_mlp: bool = True
    glu_type: str = "glu"                       # "glu": act(gate) * up   |  "swiglu": gate * sigmoid(alpha * gate) * (up + bias)
    hidden_act_scaling_factor: float = 1.0      # alpha above (GPT-OSS 1.702)
    hidden_act_bias: float = 0.0                # added to ``up`` (GPT-OSS 1.0)
    gate_clamp_upper_limit: Optional[float] = None
    gate_clamp_lower_limit: Optional[float] = None
    up_clamp_upper_limit: Optional[float] = None
    up_clamp_lower_limit: Optional[float] = None
    early_expert_affinity_modulation: bool = False
    normalize_top_k_affinities: bool = True
    is_hidden_dim_shuffled: bool = False        # MXFP4 layout flags: weights are un-shuffled at load on B200
    is_intermediate_dim_shuffled: bool = False
    use_index_calc_kernel: bool = False         # NKI-only
    enable_spmd_rank: bool = False              # SPMD tracing artefact

    def activation(self) -> Optional[Callable]:
        """``act_fn(h)`` on the fused ``h = [gate | up]`` expert projection (the engine's custom-activation hook) when the stock
        ``act(gate) * up`` epilogue does not cover the settings."""
        plain = (self.glu_type == "glu" and self.hidden_act_scaling_factor == 1.0 and self.hidden_act_bias == 0.0
                 and all(v is None for v in (self.gate_clamp_upper_limit, self.gate_clamp_lower_limit, self.up_clamp_upper_limit,
                                             self.up_clamp_lower_limit)))
        if plain:
            return None
        c = self

        def act_fn(h):
            gate, up = h.chunk(2, dim=-1)
            if c.gate_clamp_upper_limit is not None or c.gate_clamp_lower_limit is not None:
                gate = gate.clamp(min=c.gate_clamp_lower_limit, max=c.gate_clamp_upper_limit)
            if c.up_clamp_upper_limit is not None or c.up_clamp_lower_limit is not None:
                up = up.clamp(min=c.up_clamp_lower_limit, max=c.up_clamp_upper_limit)
            if c.glu_type == "swiglu":
                return (up + c.hidden_act_bias) * gate * torch.sigmoid(c.hidden_act_scaling_factor * gate)
            from .. import ops
            g = ops.activation(gate * c.hidden_act_scaling_factor, {"silu": "silu", "gelu": "gelu", "gelu_pytorch_tanh": "gelu_tanh"}[c.hidden_act])
            return g * (up + c.hidden_act_bias)
        return act_fn


@dataclass
class BlockwiseMatmulConfig:
    block_size: int = 512
    use_block_parallel: bool = False
    block_sharding_strategy: str = "hi_lo"
    skip_dma_token: bool = False
    skip_dma_weight: bool = False
    parallelize_token_to_block_mapping: bool = True
    logical_nc_config: int = 1
    use_shard_on_intermediate_dynamic_while: bool = False
    use_shard_on_block_dynamic_while: bool = False

    @classmethod
    def from_kwargs(cls, **kw):
        known = {f for f in cls.__dataclass_fields__}
        return cls(**{k: v for k, v in kw.items() if k in known})


@dataclass
class MoEFusedTKGConfig:
    quantized: bool = False
    moe_fused_kernel_enabled: Optional[bool] = None
    router_topk_kernel_enabled: Optional[bool] = None
    expert_mlp_kernel_enabled: Optional[bool] = None
    shared_mlp_kernel_enabled: Optional[bool] = None
    norm_topk_prob: bool = True
    is_mxfp4_compute: bool = False
    router_mm_dtype: torch.dtype = torch.float32

    def decode_kernel_allowed(self) -> bool:
        return self.moe_fused_kernel_enabled is not False and self.expert_mlp_kernel_enabled is not False


class ExpertMLPsV2(ExpertMLPs):
    """``ExpertMLPs`` built from the v2 config objects (reference ExpertMLPsV2(routed_experts_mlp_config, blockwise_matmul_config, ...))."""

    def __init__(self, routed_experts_mlp_config: RoutedExpertsMLPOpsConfig, blockwise_matmul_config: Optional[BlockwiseMatmulConfig] = None,
                 dtype=torch.float32, device=None, tkg_config: Optional[MoEFusedTKGConfig] = None, **groups):
        c = routed_experts_mlp_config
        if not c.glu_mlp:
            raise NotImplementedError("non-gated routed experts")
        super().__init__(c.num_experts, c.hidden_size, c.intermediate_size, c.hidden_act, dtype, c.bias, device,
                         ep_group=groups.get("ep_group"), moe_tp_group=groups.get("moe_tp_group"), act_fn=c.activation())
        self.routed_experts_mlp_config = c
        self.blockwise_matmul_config = blockwise_matmul_config or BlockwiseMatmulConfig()
        self.tkg_config = tkg_config or MoEFusedTKGConfig()

    def forward(self, x2, topk_w, topk_i, scale_input: bool = False):
        if self.tkg_config.decode_kernel_allowed():
            return super().forward(x2, topk_w, topk_i, scale_input)
        from ..ops import ref            # kernels switched off: batched-GEMM reference path
        return ref.moe_experts(x2, self.gate_up_proj, self.down_proj, topk_w, topk_i, self.act, self.expert_offset, self.gate_up_bias,
                               self.down_bias, self.act_fn, scale_input)


def initialize_moe_process_group(config, enabled_hybrid_sharding: bool = False):
    """Build the TP x EP groups of the MoE layers (reference moe_v2.py:130-161).  With hybrid sharding the reference keeps separate
    prefill / decode meshes; one process per GPU runs both phases on ONE mesh, so the decode degrees win."""
    from ..parallel import state
    nc = config.neuron_config
    tp, ep = nc.moe_tp_degree or nc.tp_degree, nc.moe_ep_degree or 1
    hs = getattr(nc, "hybrid_sharding_config", None)
    if enabled_hybrid_sharding and hs is not None:
        tp, ep = hs.moe_tkg_tp_degree or tp, hs.moe_tkg_ep_degree or ep
    if hasattr(state, "initialize_moe_groups"):
        state.initialize_moe_groups(tp, ep)
    return state.get_moe_tp_group(), state.get_expert_model_parallel_group()


def initialize_moe_module(config, *a, **kw) -> MoE:
    """v2 factory: same result as ``modules.moe.initialize_moe_module``, but the routed experts are an ``ExpertMLPsV2`` carrying the
    config objects (so per-model code can inspect / override them the way reference models do)."""
    moe = _initialize_v1(config, *a, **kw)
    nc = config.neuron_config
    e = moe.expert_mlps
    rc = RoutedExpertsMLPOpsConfig(num_experts=e.num_experts, hidden_size=config.hidden_size, intermediate_size=e.I_local * e.tp_group.size,
                                   top_k=config.num_experts_per_tok, hidden_act=getattr(config, "hidden_act", "silu"),
                                   normalize_top_k_affinities=getattr(nc, "normalize_top_k_affinities", True),
                                   early_expert_affinity_modulation=moe.early_affinity_modulation)
    e.routed_experts_mlp_config = rc
    bw = getattr(nc, "blockwise_matmul_config", None)
    e.blockwise_matmul_config = bw if isinstance(bw, BlockwiseMatmulConfig) else BlockwiseMatmulConfig.from_kwargs(**(bw or {}))
    e.tkg_config = MoEFusedTKGConfig(quantized=bool(getattr(nc, "quantized", False)),
                                     moe_fused_kernel_enabled=getattr(nc, "moe_fused_nki_kernel_enabled", None),
                                     expert_mlp_kernel_enabled=getattr(nc, "expert_mlp_nki_kernel_enabled", None),
                                     norm_topk_prob=rc.normalize_top_k_affinities)
    return moe
