"""DeepSeek-V3 / R1 (reference models/deepseek/modeling_deepseek.py:1-339 + rope_util.py, full model in
contrib/models/DeepSeek-V3): multi-head latent attention (low-rank q and kv projections, decoupled RoPE key shared by all
heads), sigmoid group-limited top-k router with a score-correction bias, routed + shared experts, leading dense layers.

Attention here is the *decompressed* MLA formulation (K = [k_nope | k_rope], V per head in the cache), which is exact and
matches Hugging Face; the weight-absorbed form that caches ``(k_pe, compressed_kv)`` (reference :79-326) is a memory
optimisation of the same function and is tracked as follow-up work in DESIGN.md."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops
from ...config import MoENeuronConfig
from ...modules.kvcache import KVCacheManager
from ...modules.mlp import GatedMLP
from ...modules.moe import ExpertMLPs, MoE, SharedExperts
from ...modules.norm import RMSNorm
from ...modules.rope import RotaryEmbedding
from ...parallel.layers import ColumnParallelLinear, ParallelEmbedding, RowParallelLinear
from ..llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaForCausalLM, rope_scaling_of, rope_theta_of
from ..model_base import DecoderLayer, NeuronBaseModel
from ..state_dict_utils import convert_moe_experts


class DeepseekInferenceConfig(LlamaInferenceConfig):
    def get_required_attributes(self):
        return ["hidden_size", "num_attention_heads", "num_hidden_layers", "vocab_size", "kv_lora_rank", "qk_nope_head_dim",
                "qk_rope_head_dim", "v_head_dim", "n_routed_experts", "num_experts_per_tok"]

    def add_derived_config(self):
        self.num_cores_per_group = 1
        self.qk_head_dim = self.qk_nope_head_dim + self.qk_rope_head_dim
        self.head_dim = self.qk_head_dim
        if not hasattr(self, "num_key_value_heads") or self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        if not hasattr(self, "hidden_act"):
            self.hidden_act = "silu"

    @classmethod
    def get_neuron_config_cls(cls):
        return MoENeuronConfig


def _yarn_mscale(scale, m):
    return 1.0 if scale <= 1 else 0.1 * m * math.log(scale) + 1.0


class DeepseekMLAttention(nn.Module):
    def __init__(self, config, layer_idx, rotary, device=None):
        super().__init__()
        nc = config.neuron_config
        dt = nc.torch_dtype
        from ...parallel.state import get_tensor_model_parallel_group
        self.tp_group = get_tensor_model_parallel_group()
        tp = self.tp_group.size
        self.layer_idx = layer_idx
        self.H = config.num_attention_heads // tp
        self.n_kv = self.n_q = self.H
        self.nope, self.rope_d, self.vd = config.qk_nope_head_dim, config.qk_rope_head_dim, config.v_head_dim
        self.head_dim = self.nope + self.rope_d
        self.q_lora = getattr(config, "q_lora_rank", None)
        Hs = config.hidden_size
        if self.q_lora:
            self.q_a_proj = nn.Linear(Hs, self.q_lora, bias=False, dtype=dt, device=device)
            self.q_a_layernorm = RMSNorm(self.q_lora, config.rms_norm_eps, dt, device=device)
            self.q_b_proj = ColumnParallelLinear(self.q_lora, config.num_attention_heads * self.head_dim, bias=False,
                                                 gather_output=False, dtype=dt, device=device)
        else:
            self.q_proj = ColumnParallelLinear(Hs, config.num_attention_heads * self.head_dim, bias=False, gather_output=False,
                                               dtype=dt, device=device)
        self.kv_a_proj_with_mqa = nn.Linear(Hs, config.kv_lora_rank + self.rope_d, bias=False, dtype=dt, device=device)
        self.kv_a_layernorm = RMSNorm(config.kv_lora_rank, config.rms_norm_eps, dt, device=device)
        self.kv_b_proj = ColumnParallelLinear(config.kv_lora_rank, config.num_attention_heads * (self.nope + self.vd), bias=False,
                                              gather_output=False, dtype=dt, device=device)
        self.o_proj = RowParallelLinear(config.num_attention_heads * self.vd, Hs, bias=False, dtype=dt, device=device)
        for p in list(self.kv_a_proj_with_mqa.parameters()) + (list(self.q_a_proj.parameters()) if self.q_lora else []):
            p.requires_grad_(False)
        self.kv_lora = config.kv_lora_rank
        self.rotary_emb = rotary
        self.scale = self.head_dim ** -0.5
        rs = rope_scaling_of(config)
        if rs and rs.get("rope_type", rs.get("type", "default")) != "default" and rs.get("mscale_all_dim", 0):
            m = _yarn_mscale(rs["factor"], rs["mscale_all_dim"])
            self.scale *= m * m
        self.sliding_window = None

    def forward(self, h, meta, kv_mgr, norm_weight=None, norm_eps=1e-6, norm_offset=0.0, residual=None, lora=None):
        B, T, _ = h.shape
        x = ops.rmsnorm(h, norm_weight, norm_eps, norm_offset) if norm_weight is not None else h
        if self.q_lora:
            q = self.q_b_proj(self.q_a_layernorm(nn.functional.linear(x, self.q_a_proj.weight)))
        else:
            q = self.q_proj(x)
        q = q.view(B, T, self.H, self.head_dim)
        q_nope, q_rot = q.split([self.nope, self.rope_d], -1)
        ckv = nn.functional.linear(x, self.kv_a_proj_with_mqa.weight)
        c, k_rot = ckv.split([self.kv_lora, self.rope_d], -1)
        kv = self.kv_b_proj(self.kv_a_layernorm(c)).view(B, T, self.H, self.nope + self.vd)
        k_nope, v = kv.split([self.nope, self.vd], -1)
        key = id(self.rotary_emb)
        if key not in meta.rope_cache:
            meta.rope_cache[key] = self.rotary_emb(meta.position_ids)
        cos, sin = meta.rope_cache[key]
        q_rot = ops.apply_rope(q_rot, cos, sin, interleaved=True)
        k_rot = ops.apply_rope(k_rot.view(B, T, 1, self.rope_d), cos, sin, interleaved=True)
        qf = torch.cat([q_nope, q_rot], -1)
        kf = torch.cat([k_nope, k_rot.expand(B, T, self.H, self.rope_d)], -1)
        if meta.lines is None:
            meta.lines = kv_mgr.lines_for(meta.seq_ids)
        kv_mgr.update(self.layer_idx, kf, v.contiguous(), meta.seq_ids, meta.write_positions, meta.lines)
        if meta.is_prefill and not meta.has_prefix:
            o = ops.ref.attention_prefill(qf, kf, v, self.scale, True, None, None,
                                          None if self_right(meta) else meta.key_valid, None if self_right(meta) else meta.position_ids)
        else:
            kc, vc = kv_mgr.get_kv_by_layer_id(self.layer_idx)
            o = ops.ref.attention_decode(qf, kc, vc, meta.lines, meta.position_ids, self.scale)
        return self.o_proj(o.reshape(B, T, self.H * self.vd), residual)


def self_right(meta):
    return not getattr(meta, "offset_positions", False)


class DeepseekRouter(nn.Module):
    """sigmoid scores, + e_score_correction_bias for *selection only*, group-limited top-k, renormalise, scale."""

    def __init__(self, config, device=None):
        super().__init__()
        self.E, self.top_k = config.n_routed_experts, config.num_experts_per_tok
        self.n_group, self.topk_group = getattr(config, "n_group", 1) or 1, getattr(config, "topk_group", 1) or 1
        self.norm = bool(getattr(config, "norm_topk_prob", True))
        self.scaling = float(getattr(config, "routed_scaling_factor", 1.0))
        self.linear_router = nn.Linear(config.hidden_size, self.E, bias=False, dtype=torch.float32, device=device)
        self.linear_router.weight.requires_grad_(False)
        self.register_buffer("e_score_correction_bias", torch.zeros(self.E, dtype=torch.float32, device=device))

    def forward(self, x):
        logits = nn.functional.linear(x.float(), self.linear_router.weight)
        s = logits.sigmoid()
        choice = s + self.e_score_correction_bias
        N = s.shape[0]
        g = choice.view(N, self.n_group, self.E // self.n_group)
        gscore = g.topk(min(2, g.shape[-1]), -1)[0].sum(-1)
        gidx = gscore.topk(self.topk_group, -1)[1]
        gmask = torch.zeros_like(gscore).scatter_(1, gidx, 1.0)
        smask = gmask.unsqueeze(-1).expand(N, self.n_group, self.E // self.n_group).reshape(N, self.E).bool()
        idx = choice.masked_fill(~smask, 0.0).topk(self.top_k, -1)[1]
        w = s.gather(1, idx)
        if self.norm:
            w = w / (w.sum(-1, keepdim=True) + 1e-20)
        return logits, w * self.scaling, idx


class NeuronDeepseekModel(NeuronBaseModel):
    graph_safe = False
    router_cls = DeepseekRouter

    def setup_attr_for_model(self, config):
        nc = config.neuron_config
        self.tp_degree, self.hidden_size = nc.tp_degree, config.hidden_size
        self.num_attention_heads = self.num_key_value_heads = config.num_attention_heads
        self.max_batch_size, self.buckets = nc.max_batch_size, nc.buckets

    def kv_heads_per_rank(self):
        return self.layers[0].self_attn.H

    def kv_head_dim(self):
        return self.layers[0].self_attn.head_dim

    def init_inference_optimization(self, config):
        from ...modules.sampling import Sampler
        nc = self.neuron_config
        if self.on_device_sampling:
            self.sampler = Sampler(nc, self.tp_group, vocab_shard=self.lm_head_is_sharded())
        a = self.layers[0].self_attn
        self.kv_mgr = KVCacheManager(len(self.layers), a.H, a.head_dim, nc.max_length + nc.speculation_length,
                                     nc.kv_cache_batch_size + nc.kv_cache_padding_size, nc.attention_dtype or nc.torch_dtype,
                                     self.device_, v_head_dim=a.vd)

    def init_model(self, config):
        nc = config.neuron_config
        dev, dt = self.device_, nc.torch_dtype
        self.embed_tokens = ParallelEmbedding(config.vocab_size, config.hidden_size, None, dtype=dt, device=dev,
                                              shard_across_embedding=not nc.vocab_parallel, pad=True,
                                              tensor_model_parallel_group=self.tp_group)
        rotary = RotaryEmbedding(config.qk_rope_head_dim, max(getattr(config, "max_position_embeddings", 4096), nc.seq_len),
                                 rope_theta_of(config), rope_scaling_of(config), device=dev)
        layers = []
        first_dense = getattr(config, "first_k_dense_replace", 0)
        for i in range(config.num_hidden_layers):
            attn = DeepseekMLAttention(config, i, rotary, dev)
            if i >= first_dense and i % getattr(config, "moe_layer_freq", 1) == 0:
                experts = ExpertMLPs(config.n_routed_experts, config.hidden_size, config.moe_intermediate_size, config.hidden_act, dt,
                                     device=dev)
                shared = SharedExperts(config.hidden_size, config.moe_intermediate_size * (getattr(config, "n_shared_experts", 1) or 1),
                                       config.hidden_act, dt, dev) if getattr(config, "n_shared_experts", 0) else None
                mlp = MoE(self.router_cls(config, dev), experts, shared)
            else:
                mlp = GatedMLP(config.hidden_size, config.intermediate_size, config.hidden_act, dt, device=dev)
            layers.append(DecoderLayer(attn, mlp, RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=dev),
                                       RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=dev), i))
        self.layers = nn.ModuleList(layers)
        self.norm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=dev)
        self.lm_head = ColumnParallelLinear(config.hidden_size, config.vocab_size, bias=False, gather_output=False, dtype=dt,
                                            device=dev, pad=True, tensor_model_parallel_group=self.tp_group)


class NeuronDeepseekForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronDeepseekModel

    @classmethod
    def get_config_cls(cls):
        return DeepseekInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict, config):
        sd = dict(state_dict)
        L = config.num_hidden_layers
        for i in range(L):
            m = f"layers.{i}.mlp."
            for sfx in ("weight",):
                ks = [f"{m}gate_proj.{sfx}", f"{m}up_proj.{sfx}"]
                if all(k in sd for k in ks):
                    sd[f"{m}gate_up_proj.{sfx}"] = torch.cat([sd.pop(k) for k in ks], 0)
                ks = [f"{m}shared_experts.gate_proj.{sfx}", f"{m}shared_experts.up_proj.{sfx}"]
                if all(k in sd for k in ks):
                    sd[f"{m}shared_experts.gate_up_proj.{sfx}"] = torch.cat([sd.pop(k) for k in ks], 0)
            if f"{m}gate.e_score_correction_bias" in sd:
                sd[f"{m}router.e_score_correction_bias"] = sd.pop(f"{m}gate.e_score_correction_bias").float()
        sd = convert_moe_experts(sd, L, config.n_routed_experts, moe_prefixes=("mlp",), gate_names=("gate",),
                                 w_names=("gate_proj", "up_proj", "down_proj"))
        return {k: v for k, v in sd.items() if "rotary_emb" not in k}
