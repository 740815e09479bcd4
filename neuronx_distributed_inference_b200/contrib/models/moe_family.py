"""Mixture-of-experts community families on the engine's MoE blocks (router + EP x TP experts + optional shared expert):

* **Qwen2-MoE / Qwen1.5-MoE** — Qwen2 attention, top-k softmax router, shared expert scaled by ``sigmoid(w_g . x)``.
* **OLMoE** — Llama block with q/k RMSNorm over the whole projection, 64-expert style router without renormalisation.
* **EXAONE-4** (dense, listed here because it shares the post-norm block) — post-norm residuals, per-head q/k RMSNorm, hybrid
  sliding / global layers with rotary only on the sliding ones.
* **GraniteMoE** — Granite multipliers around a top-k MoE whose experts ship as fused ``input_linear`` / ``output_linear``.
* **Phi-3.5-MoE** — LayerNorm block, biased attention / head, SparseMixer top-2 routing (jitter-thresholded softmax per pick).
* **GLM-4.5 (``glm4_moe``) / dots.llm1 (``dots1``)** — GQA attention (partial rotary / per-head q,k RMSNorm) in front of the
  DeepSeek-V3 MoE block (sigmoid scores, selection-only correction bias, group-limited top-k, shared experts, dense first layers).
* **DeepSeek-V2 / V2-Lite** — the MLA attention of the DeepSeek-V3 model with the V2 router (softmax scores, greedy or
  group-limited-greedy selection by group maximum, no renormalisation).
* **Trinity (Arcee AFMoE)** — sandwich norms, per-head q/k RMSNorm, RoPE only on the sliding-window layers, a sigmoid OUTPUT GATE on
  the attention (``o * sigmoid(W_g x)`` before o_proj), dense first layers then sigmoid-routed MoE with a selection-only expert bias
  and shared experts, muP embedding scale.
* **ERNIE-4.5-MoE** — interleaved rotary, softmax router whose correction bias steers selection only, shared experts, MoE layer window.
reference ports: contrib/models/{EXAONE-4.0-1.2B, Phi-3.5-MoE-instruct}/src and the MoE glue of modules/moe_v2.py."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ...config import MoENeuronConfig
from ...models.deepseek.modeling_deepseek import NeuronDeepseekForCausalLM, NeuronDeepseekModel
from ...models.llama.modeling_llama import (LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaMLP,
                                            NeuronLlamaModel)
from ...models.model_base import DecoderLayer
from ...models.qwen2.modeling_qwen2 import NeuronQwen2Attention
from ...models.state_dict_utils import convert_moe_experts, fuse_qkv_and_gate_up
from ...modules.attention import AttentionBase
from ...modules.mlp import GatedMLP
from ...modules.moe import initialize_moe_module
from ...modules.norm import RMSNorm
from .classic_family import NeuronClassicModel, _ClassicCausalLM
from .llama_family import Olmo2Attention


class _MoeConfig(LlamaInferenceConfig):
    @classmethod
    def get_neuron_config_cls(cls):
        return MoENeuronConfig


def _is_moe_layer(config, i):
    if i in (getattr(config, "mlp_only_layers", None) or []):
        return False
    step = getattr(config, "decoder_sparse_step", 1) or 1
    return getattr(config, "num_experts", 0) > 0 and (i + 1) % step == 0


# ---- Qwen2-MoE ----------------------------------------------------------------------------------------------------------------
class NeuronQwen2MoeModel(NeuronLlamaModel):
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        nc = config.neuron_config
        dt = nc.torch_dtype
        attn = NeuronQwen2Attention(config, i, rotary, device=device)
        if _is_moe_layer(config, i):
            mlp = initialize_moe_module(config, device=device, intermediate_size=config.moe_intermediate_size,
                                        normalize=bool(getattr(config, "norm_topk_prob", False)))
            from ...modules.moe import SharedExperts
            mlp.shared_experts = SharedExperts(config.hidden_size, config.shared_expert_intermediate_size, config.hidden_act, dt, device)
            mlp.shared_expert_gate = nn.Linear(config.hidden_size, 1, bias=False, dtype=dt, device=device)
            mlp.shared_expert_gate.weight.requires_grad_(False)
        else:
            mlp = NeuronLlamaMLP(config, device=device)
        return DecoderLayer(attn, mlp, RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device), i, mlp_is_moe=_is_moe_layer(config, i))


class NeuronQwen2MoeForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronQwen2MoeModel

    @classmethod
    def get_config_cls(cls):
        return _MoeConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        moe_layers = [i for i in range(config.num_hidden_layers) if _is_moe_layer(config, i)]
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=True)
        sd = convert_moe_experts(sd, config.num_hidden_layers, config.num_experts, moe_prefixes=("mlp",), gate_names=("gate",),
                                 w_names=("gate_proj", "up_proj", "down_proj"), layers=moe_layers)
        out = {}
        for k, v in sd.items():
            k = k.replace(".mlp.shared_expert.", ".mlp.shared_experts.")
            out[k] = v
        for i in moe_layers:
            g, u = f"layers.{i}.mlp.shared_experts.gate_proj.weight", f"layers.{i}.mlp.shared_experts.up_proj.weight"
            if g in out:
                out[f"layers.{i}.mlp.shared_experts.gate_up_proj.weight"] = torch.cat([out.pop(g), out.pop(u)], 0)
        return out


# ---- OLMoE ------------------------------------------------------------------------------------------------------------------------
class NeuronOlmoeModel(NeuronLlamaModel):
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        dt = config.neuron_config.torch_dtype
        attn = Olmo2Attention(config, i, rotary, device)
        moe = initialize_moe_module(config, device=device, intermediate_size=config.intermediate_size,
                                    normalize=bool(getattr(config, "norm_topk_prob", False)))
        return DecoderLayer(attn, moe, RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device), i, mlp_is_moe=True)


class NeuronOlmoeForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronOlmoeModel

    @classmethod
    def get_config_cls(cls):
        return _MoeConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=True)
        sd = convert_moe_experts(sd, config.num_hidden_layers, config.num_experts, moe_prefixes=("mlp",), gate_names=("gate",),
                                 w_names=("gate_proj", "up_proj", "down_proj"))
        return {k.replace(".self_attn.q_norm.weight", ".self_attn.q_norm").replace(".self_attn.k_norm.weight", ".self_attn.k_norm"): v
                for k, v in sd.items()}


# ---- EXAONE-4 ----------------------------------------------------------------------------------------------------------------------
class Exaone4DecoderLayer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        lt = getattr(config, "layer_types", None)
        hybrid = getattr(config, "sliding_window", None) is not None
        sliding = bool(lt and lt[i] == "sliding_attention")
        self.self_attn = AttentionBase(config, hidden_size=config.hidden_size, num_attention_heads=config.num_attention_heads,
                                       num_key_value_heads=config.num_key_value_heads, head_dim=config.head_dim, rotary_emb=rotary,
                                       qk_norm="rms_pre_rope", qk_norm_eps=config.rms_norm_eps, use_rope=(not hybrid) or sliding,
                                       sliding_window=config.sliding_window if (hybrid and sliding) else None, layer_idx=i,
                                       rms_norm_eps=config.rms_norm_eps, device=device)
        self.mlp = GatedMLP(config.hidden_size, config.intermediate_size, config.hidden_act, dt, device=device)
        self.post_attention_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.post_feedforward_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.layer_idx = i

    def forward(self, h, meta, kv_mgr, lora=None):
        h = h + self.post_attention_layernorm(self.self_attn(h, meta, kv_mgr))
        return h + self.post_feedforward_layernorm(self.mlp(h))


class NeuronExaone4Model(NeuronLlamaModel):
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        return Exaone4DecoderLayer(config, i, rotary, device)


class NeuronExaone4ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronExaone4Model

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers)
        return {k.replace("self_attn.q_norm.", "self_attn.q_layernorm.").replace("self_attn.k_norm.", "self_attn.k_layernorm."): v
                for k, v in sd.items()}


# ---- GraniteMoE -------------------------------------------------------------------------------------------------------------------
class NeuronGraniteMoeModel(NeuronLlamaModel):
    graph_safe = False

    def init_model(self, config):
        super().init_model(config)
        self.embed_scale = float(getattr(config, "embedding_multiplier", 1.0))

    def make_layer(self, config, i, rotary, device):
        dt = config.neuron_config.torch_dtype
        attn = self.attention_cls(config, i, rotary, device=device, softmax_scale=float(getattr(config, "attention_multiplier", None)
                                                                                        or config.head_dim ** -0.5))
        moe = initialize_moe_module(config, device=device, normalize=True)      # softmax over the selected logits == renormalised top-k
        return DecoderLayer(attn, moe, RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device), i, mlp_is_moe=True)


class NeuronGraniteMoeForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronGraniteMoeModel

    @classmethod
    def get_config_cls(cls):
        return _MoeConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        """The residual multiplier scales linear outputs (o_proj, expert down projections) and the logits scaling divides the head:
        both are folded into the weights (same trick as dense Granite)."""
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
        rm, ls = float(getattr(config, "residual_multiplier", 1.0)), float(getattr(config, "logits_scaling", 1.0))
        out = {}
        for k, v in sd.items():
            if k.endswith(".self_attn.o_proj.weight"):
                v = (v.float() * rm).to(v.dtype)
            elif k.endswith(".block_sparse_moe.output_linear.weight"):
                k, v = k.replace(".block_sparse_moe.output_linear.weight", ".mlp.expert_mlps.down_proj"), (v.float() * rm).to(v.dtype)
            elif k.endswith(".block_sparse_moe.input_linear.weight"):
                k = k.replace(".block_sparse_moe.input_linear.weight", ".mlp.expert_mlps.gate_up_proj")
            elif k.endswith(".block_sparse_moe.router.layer.weight"):
                k, v = k.replace(".block_sparse_moe.router.layer.weight", ".mlp.router.linear_router.weight"), v.float()
            out[k] = v
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        out["lm_head.weight"] = (out["lm_head.weight"].float() / ls).to(out["lm_head.weight"].dtype)
        return out

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        pass


# ---- Phi-3.5-MoE ------------------------------------------------------------------------------------------------------------------
class SparseMixerRouter(nn.Module):
    """Inference form of SparseMixer (arXiv 2409.12136): pick the arg-max expert, weight it by a softmax over the logits that lie
    within a relative ``2 * jitter`` band of the maximum; remove it and repeat once for the second expert."""

    def __init__(self, num_experts, hidden_size, jitter_eps, device=None):
        super().__init__()
        self.num_experts, self.top_k, self.jitter_eps = num_experts, 2, float(jitter_eps)
        self.linear_router = nn.Linear(hidden_size, num_experts, bias=False, dtype=torch.float32, device=device)
        self.linear_router.weight.requires_grad_(False)

    def _pick(self, scores, pool):
        top, idx = pool.max(-1, keepdim=True)
        factor = scores.abs().clamp(min=top)
        gates = torch.softmax(pool.masked_fill(((top - scores) / factor) > 2 * self.jitter_eps, float("-inf")), -1)
        return gates.gather(-1, idx), idx

    def forward(self, x):
        scores = nn.functional.linear(x.float(), self.linear_router.weight)
        w1, i1 = self._pick(scores, scores)
        w2, i2 = self._pick(scores, scores.scatter(-1, i1, float("-inf")))
        return scores, torch.cat([w1, w2], -1), torch.cat([i1, i2], -1)


def _phimoe_block(config, device):
    moe = initialize_moe_module(config, device=device, normalize=False)
    moe.router = SparseMixerRouter(config.num_local_experts, config.hidden_size, getattr(config, "router_jitter_noise", 0.01), device)
    return moe


class _PhimoeConfig(_MoeConfig):
    pass


class NeuronPhimoeModel(NeuronClassicModel):
    def init_model(self, config):
        self.lm_head_bias = bool(getattr(config, "lm_head_bias", False))
        super().init_model(config)

    def layer_spec(self, config, i):
        b = bool(getattr(config, "attention_bias", False))
        return dict(parallel=False, norm_bias=True, mlp=_phimoe_block, act=config.hidden_act, qkv_bias=b, o_bias=b, mlp_bias=False,
                    sliding_window=getattr(config, "sliding_window", None))


class NeuronPhimoeForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronPhimoeModel

    @classmethod
    def get_config_cls(cls):
        return _PhimoeConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
        return convert_moe_experts(sd, config.num_hidden_layers, config.num_local_experts, moe_prefixes=("mlp", "block_sparse_moe"),
                                   gate_names=("router", "gate"), w_names=("w1", "w3", "w2"))


# ---- GLM-4.5-MoE / dots.llm1: GQA attention + DeepSeek-V3 MoE ----------------------------------------------------------------------
def _deepseek_moe(config, device):
    from ...models.deepseek.modeling_deepseek import DeepseekRouter
    from ...modules.moe import ExpertMLPs, MoE, SharedExperts
    dt = config.neuron_config.torch_dtype
    experts = ExpertMLPs(config.n_routed_experts, config.hidden_size, config.moe_intermediate_size, config.hidden_act, dt, device=device)
    n_sh = getattr(config, "n_shared_experts", 0) or 0
    shared = SharedExperts(config.hidden_size, config.moe_intermediate_size * n_sh, config.hidden_act, dt, device) if n_sh else None
    return MoE(DeepseekRouter(config, device), experts, shared)


class _Glm4MoeAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        if getattr(config, "use_qk_norm", False):
            over = dict(over, qk_norm="rms_pre_rope", qk_norm_eps=config.rms_norm_eps)
        super().__init__(config, layer_idx, rotary_emb, device=device, qkv_bias=bool(getattr(config, "attention_bias", False)),
                         o_bias=False, **over)


class _Dots1Attention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        lt = getattr(config, "layer_types", None)
        sw = getattr(config, "sliding_window", None) if (lt and lt[layer_idx] == "sliding_attention") else None
        super().__init__(config, layer_idx, rotary_emb, device=device, qk_norm="rms_pre_rope", qk_norm_eps=config.rms_norm_eps,
                         sliding_window=sw, **over)


class NeuronGlm4MoeModel(NeuronLlamaModel):
    attention_cls = _Glm4MoeAttention
    graph_safe = False

    def make_rotary(self, config, device):
        from ...models.llama.modeling_llama import rope_scaling_of, rope_theta_of
        from ...modules.rope import RotaryEmbedding
        rp = getattr(config, "rope_parameters", None) or {}
        frac = getattr(config, "partial_rotary_factor", None) or (rp.get("partial_rotary_factor") if isinstance(rp, dict) else None) or 1.0
        return RotaryEmbedding(int(config.head_dim * float(frac)), max(config.max_position_embeddings, config.neuron_config.seq_len),
                               rope_theta_of(config), rope_scaling_of(config), device=device)

    def make_layer(self, config, i, rotary, device):
        dt = config.neuron_config.torch_dtype
        moe = i >= getattr(config, "first_k_dense_replace", 0)
        mlp = _deepseek_moe(config, device) if moe else NeuronLlamaMLP(config, device=device)
        return DecoderLayer(self.attention_cls(config, i, rotary, device=device), mlp,
                            RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device), i, mlp_is_moe=moe)


class NeuronGlm4MoeForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronGlm4MoeModel

    @classmethod
    def get_config_cls(cls):
        return _MoeConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        from ...models.deepseek.modeling_deepseek import NeuronDeepseekForCausalLM
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
        sd = NeuronDeepseekForCausalLM.convert_hf_to_neuron_state_dict(sd, config)
        return {k.replace("self_attn.q_norm.", "self_attn.q_layernorm.").replace("self_attn.k_norm.", "self_attn.k_layernorm."): v
                for k, v in sd.items()}


class NeuronDots1Model(NeuronGlm4MoeModel):
    attention_cls = _Dots1Attention


class NeuronDots1ForCausalLM(NeuronGlm4MoeForCausalLM):
    _model_cls = NeuronDots1Model


# ---- Trinity (AFMoE) --------------------------------------------------------------------------------------------------------------
class _AfmoeAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        lt = getattr(config, "layer_types", None)
        local = bool(lt and lt[layer_idx] == "sliding_attention")
        super().__init__(config, layer_idx, rotary_emb, device=device, qk_norm="rms_pre_rope", qk_norm_eps=config.rms_norm_eps,
                         use_rope=local, sliding_window=getattr(config, "sliding_window", None) if local else None, **over)
        dt, D = config.neuron_config.torch_dtype, self.head_dim
        plan = self.qkv_proj.plan
        self.gate_weight = nn.Parameter(torch.zeros(self.n_q * D, config.hidden_size, dtype=dt, device=device), requires_grad=False)
        from ...modules.gqa import _gather_heads
        self.gate_weight.shard_fn = lambda full, rank: _gather_heads(full, plan.q_idx[rank], D, 0)
        self.gate_weight.partition_dim, self.gate_weight.tp_group = 0, self.tp_group
        self._gate = None

    def forward(self, hidden, meta, kv_mgr, norm_weight=None, norm_eps=None, norm_offset=0.0, residual=None, lora=None):
        xn = ops.rmsnorm(hidden, norm_weight, norm_eps if norm_eps is not None else self.rms_norm_eps, norm_offset) \
            if norm_weight is not None else hidden
        self._gate = torch.sigmoid(ops.linear(xn, self.gate_weight))
        return super().forward(hidden, meta, kv_mgr, norm_weight, norm_eps, norm_offset, residual, lora)

    def _finish(self, o, residual, lora, meta):
        return super()._finish(o * self._gate.to(o.dtype), residual, lora, meta)


class AfmoeDecoderLayer(nn.Module):
    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        self.self_attn = _AfmoeAttention(config, i, rotary, device=device)
        self.mlp_is_moe = i >= getattr(config, "num_dense_layers", 0)
        self.mlp = _deepseek_moe(config, device) if self.mlp_is_moe else NeuronLlamaMLP(config, device=device)
        mk = lambda: RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)   # noqa: E731
        self.input_layernorm, self.post_attention_layernorm, self.pre_mlp_layernorm, self.post_mlp_layernorm = mk(), mk(), mk(), mk()
        self.layer_idx = i

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.input_layernorm
        h = h + self.post_attention_layernorm(self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon))
        n = self.pre_mlp_layernorm
        return h + self.post_mlp_layernorm(self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon))


class AfmoeInferenceConfig(_MoeConfig):
    def add_derived_config(self):
        # names the shared DeepSeek-style MoE block reads
        self.n_routed_experts, self.n_shared_experts = self.num_experts, getattr(self, "num_shared_experts", 0)
        self.routed_scaling_factor, self.norm_topk_prob, self.n_group, self.topk_group = getattr(self, "route_scale", 1.0), True, 1, 1
        super().add_derived_config()


class NeuronTrinityModel(NeuronLlamaModel):
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        return AfmoeDecoderLayer(config, i, rotary, device)

    def init_model(self, config):
        super().init_model(config)
        if getattr(config, "mup_enabled", False):
            self.embed_scale = float(config.hidden_size ** 0.5)


class NeuronTrinityForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronTrinityModel

    @classmethod
    def get_config_cls(cls):
        return AfmoeInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=True)
        for i in range(config.num_hidden_layers):
            m = f"layers.{i}.mlp."
            if m + "expert_bias" in sd:
                sd[m + "router.e_score_correction_bias"] = sd.pop(m + "expert_bias").float()
            if m + "router.gate.weight" in sd:
                sd[m + "router.linear_router.weight"] = sd.pop(m + "router.gate.weight").float()
            g, u = m + "shared_experts.gate_proj.weight", m + "shared_experts.up_proj.weight"
            if g in sd:
                sd[m + "shared_experts.gate_up_proj.weight"] = torch.cat([sd.pop(g), sd.pop(u)], 0)
            a = f"layers.{i}.self_attn."
            if a + "gate_proj.weight" in sd:
                sd[a + "gate_weight"] = sd.pop(a + "gate_proj.weight")
        sd = convert_moe_experts(sd, config.num_hidden_layers, config.num_experts, moe_prefixes=("mlp",), gate_names=(),
                                 w_names=("gate_proj", "up_proj", "down_proj"))
        return {k.replace("self_attn.q_norm.", "self_attn.q_layernorm.").replace("self_attn.k_norm.", "self_attn.k_layernorm."): v
                for k, v in sd.items()}


# ---- DeepSeek-V2 ------------------------------------------------------------------------------------------------------------------
class DeepseekV2Router(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        self.E, self.top_k = config.n_routed_experts, config.num_experts_per_tok
        self.method = getattr(config, "topk_method", "greedy")
        self.n_group, self.topk_group = getattr(config, "n_group", 1) or 1, getattr(config, "topk_group", 1) or 1
        self.norm = bool(getattr(config, "norm_topk_prob", False))
        self.scaling = float(getattr(config, "routed_scaling_factor", 1.0))
        self.linear_router = nn.Linear(config.hidden_size, self.E, bias=False, dtype=torch.float32, device=device)
        self.linear_router.weight.requires_grad_(False)

    def forward(self, x):
        logits = nn.functional.linear(x.float(), self.linear_router.weight)
        p = torch.softmax(logits, -1)
        pool = p
        if self.method == "group_limited_greedy":
            N = p.shape[0]
            gmax = p.view(N, self.n_group, -1).max(-1).values
            keep = torch.zeros_like(gmax).scatter_(1, gmax.topk(self.topk_group, -1)[1], 1.0).bool()
            pool = p.masked_fill(~keep.unsqueeze(-1).expand(N, self.n_group, self.E // self.n_group).reshape(N, self.E), 0.0)
        w, idx = pool.topk(self.top_k, -1)
        if self.norm:
            w = w / (w.sum(-1, keepdim=True) + 1e-20)
        return logits, w * self.scaling, idx


class NeuronDeepseekV2Model(NeuronDeepseekModel):
    router_cls = DeepseekV2Router


class NeuronDeepseekV2ForCausalLM(NeuronDeepseekForCausalLM):
    _model_cls = NeuronDeepseekV2Model


# ---- ERNIE-4.5-MoE ----------------------------------------------------------------------------------------------------------------
class ErnieMoeRouter(nn.Module):
    """softmax(logits); the correction bias is added for the top-k SELECTION only; selected probabilities are renormalised."""

    def __init__(self, num_experts, top_k, hidden_size, norm_min=1e-12, device=None):
        super().__init__()
        self.num_experts, self.top_k, self.norm_min = num_experts, top_k, float(norm_min)
        self.linear_router = nn.Linear(hidden_size, num_experts, bias=False, dtype=torch.float32, device=device)
        self.linear_router.weight.requires_grad_(False)
        self.register_buffer("e_score_correction_bias", torch.zeros(num_experts, dtype=torch.float32, device=device))

    def forward(self, x):
        logits = nn.functional.linear(x.float(), self.linear_router.weight)
        p = torch.softmax(logits, -1)
        idx = (p + self.e_score_correction_bias).topk(self.top_k, -1)[1]
        w = p.gather(-1, idx)
        return logits, w / w.sum(-1, keepdim=True).clamp(min=self.norm_min), idx


def _ernie_is_moe(config, i):
    return ((i + 1) % getattr(config, "moe_layer_interval", 1) == 0 and i >= getattr(config, "moe_layer_start_index", 0)
            and i <= getattr(config, "moe_layer_end_index", config.num_hidden_layers - 1))


class NeuronErnie4_5MoeModel(NeuronLlamaModel):
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        from ...modules.moe import ExpertMLPs, MoE, SharedExperts
        dt = config.neuron_config.torch_dtype
        b = bool(getattr(config, "use_bias", False))
        attn = NeuronLlamaAttention(config, i, rotary, device=device, rope_interleaved=True, qkv_bias=b, o_bias=b)
        moe = _ernie_is_moe(config, i)
        if moe:
            experts = ExpertMLPs(config.moe_num_experts, config.hidden_size, config.moe_intermediate_size, config.hidden_act, dt, device=device)
            n_sh = getattr(config, "moe_num_shared_experts", 0) or 0
            shared = SharedExperts(config.hidden_size, config.moe_intermediate_size * n_sh, config.hidden_act, dt, device) if n_sh else None
            mlp = MoE(ErnieMoeRouter(config.moe_num_experts, config.moe_k, config.hidden_size, getattr(config, "moe_norm_min", 1e-12), device),
                      experts, shared)
        else:
            mlp = NeuronLlamaMLP(config, device=device)
        return DecoderLayer(attn, mlp, RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device), i, mlp_is_moe=moe)


class NeuronErnie4_5MoeForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronErnie4_5MoeModel

    @classmethod
    def get_config_cls(cls):
        return _MoeConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=True)
        for i in range(config.num_hidden_layers):
            m = f"layers.{i}.mlp."
            b = sd.pop(m + "gate.moe_statics.e_score_correction_bias", None)
            b = sd.pop(m + "moe_statics.e_score_correction_bias", b)        # on-disk spelling (the in-memory module nests it under gate)
            if b is not None:
                sd[m + "router.e_score_correction_bias"] = b.float().reshape(-1)
            g, u = m + "shared_experts.gate_proj.weight", m + "shared_experts.up_proj.weight"
            if g in sd:
                sd[m + "shared_experts.gate_up_proj.weight"] = torch.cat([sd.pop(g), sd.pop(u)], 0)
        return convert_moe_experts(sd, config.num_hidden_layers, config.moe_num_experts, moe_prefixes=("mlp",), gate_names=("gate",),
                                   w_names=("gate_proj", "up_proj", "down_proj"))


MOE_MODEL_TYPES = {"afmoe": NeuronTrinityForCausalLM, "deepseek_v2": NeuronDeepseekV2ForCausalLM, "glm4_moe": NeuronGlm4MoeForCausalLM, "dots1": NeuronDots1ForCausalLM, "ernie4_5_moe": NeuronErnie4_5MoeForCausalLM,
                   "granitemoe": NeuronGraniteMoeForCausalLM, "phimoe": NeuronPhimoeForCausalLM,
                   "qwen2_moe": NeuronQwen2MoeForCausalLM, "olmoe": NeuronOlmoeForCausalLM, "exaone4": NeuronExaone4ForCausalLM}
