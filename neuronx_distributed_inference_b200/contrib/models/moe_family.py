"""Mixture-of-experts community families on the engine's MoE blocks (router + EP x TP experts + optional shared expert):

* **Qwen2-MoE / Qwen1.5-MoE** — Qwen2 attention, top-k softmax router, shared expert scaled by ``sigmoid(w_g . x)``.
* **OLMoE** — Llama block with q/k RMSNorm over the whole projection, 64-expert style router without renormalisation.
* **EXAONE-4** (dense, listed here because it shares the post-norm block) — post-norm residuals, per-head q/k RMSNorm, hybrid
  sliding / global layers with rotary only on the sliding ones.
reference ports: contrib/models/{EXAONE-4.0-1.2B}/src and the MoE glue of modules/moe_v2.py."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...config import MoENeuronConfig
from ...models.llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaForCausalLM, NeuronLlamaMLP, NeuronLlamaModel
from ...models.model_base import DecoderLayer
from ...models.qwen2.modeling_qwen2 import NeuronQwen2Attention
from ...models.state_dict_utils import convert_moe_experts, fuse_qkv_and_gate_up
from ...modules.attention import AttentionBase
from ...modules.mlp import GatedMLP
from ...modules.moe import initialize_moe_module
from ...modules.norm import RMSNorm
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


MOE_MODEL_TYPES = {"qwen2_moe": NeuronQwen2MoeForCausalLM, "olmoe": NeuronOlmoeForCausalLM, "exaone4": NeuronExaone4ForCausalLM}
