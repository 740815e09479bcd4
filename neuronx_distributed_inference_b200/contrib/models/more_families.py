"""More community-hub decoders, each a small delta over the Llama / classic blocks:

* **Gemma (v1)** — offset RMSNorm (1 + w), sqrt(H)-scaled embeddings, GeGLU, explicit head_dim, tied head.
* **VaultGemma** — Gemma-2 attention (soft-cap, ``query_pre_attn_scalar``, sliding / global layers) in a plain pre-norm block.
* **GLM-4-9B-chat (``glm``)** — partial (50 %) interleaved rotary, q/k/v biases, fused gate_up in the checkpoint.
* **Cohere2 / Command-R7B** — Cohere block; sliding-window layers use RoPE, full-attention layers have NO position encoding.
* **Apertus** — per-head q/k RMSNorm, non-gated MLP with the learned xIELU activation.
* **Nemotron** — LayerNorm1P (1 + w, folded into the weight at load), partial rotary, squared-ReLU non-gated MLP.
reference ports: contrib/models/{gemma-2b-it, vaultgemma-1b, glm-4-9b-chat-hf, c4ai-command-r7b-12-2024, Apertus-8B-Instruct-2509}/src."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...models.gemma3.modeling_gemma3 import Gemma3InferenceConfig, _is_sliding
from ...models.llama.modeling_llama import NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel, rope_scaling_of, rope_theta_of
from ...models.model_base import DecoderLayer
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from ...modules.mlp import PlainMLP
from ...modules.norm import RMSNorm
from ...modules.rope import RotaryEmbedding
from .classic_family import NeuronClassicModel, _ClassicCausalLM, _rename_plain_mlp


def _partial_rotary(config, default=1.0):
    rp = getattr(config, "rope_parameters", None)
    f = getattr(config, "partial_rotary_factor", None)
    if f is None and isinstance(rp, dict):
        f = rp.get("partial_rotary_factor")
    return float(f if f is not None else default)


# ---------------------------------------------------------------------------------------------------------- Gemma (v1)
class NeuronGemmaModel(NeuronLlamaModel):
    def _norm(self, config, device):
        return RMSNorm(config.hidden_size, config.rms_norm_eps, config.neuron_config.torch_dtype, offset=1.0, device=device)

    def make_layer(self, config, i, rotary, device):
        return DecoderLayer(self.attention_cls(config, i, rotary, device=device), self.mlp_cls(config, device=device),
                            self._norm(config, device), self._norm(config, device), i)

    def init_model(self, config):
        super().init_model(config)
        self.norm = self._norm(config, self.device_)
        self.embed_scale = float(torch.tensor(config.hidden_size ** 0.5, dtype=config.neuron_config.torch_dtype))


class NeuronGemmaForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronGemmaModel

    @classmethod
    def get_config_cls(cls):
        return Gemma3InferenceConfig


# ---------------------------------------------------------------------------------------------------------- VaultGemma
class _VaultGemmaAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        sliding = _is_sliding(config, layer_idx) if getattr(config, "layer_types", None) else False
        scalar = getattr(config, "query_pre_attn_scalar", config.head_dim)
        super().__init__(config, layer_idx, rotary_emb, device=device, softmax_scale=1.0 / math.sqrt(scalar),
                         logit_softcap=getattr(config, "attn_logit_softcapping", None),
                         sliding_window=config.sliding_window if sliding else None, **over)


class NeuronVaultGemmaModel(NeuronGemmaModel):
    attention_cls = _VaultGemmaAttention
    graph_safe = False

    def init_model(self, config):
        super().init_model(config)
        self.final_logit_softcap = getattr(config, "final_logit_softcapping", None)


class NeuronVaultGemmaForCausalLM(NeuronGemmaForCausalLM):
    _model_cls = NeuronVaultGemmaModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers)
        return {k.replace(".pre_feedforward_layernorm.", ".post_attention_layernorm."): v for k, v in sd.items()}


# ---------------------------------------------------------------------------------------------------------- GLM (glm-4-9b-chat-hf)
class _GlmAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, rope_interleaved=True,
                         qkv_bias=bool(getattr(config, "attention_bias", True)), o_bias=False, **over)


class NeuronGlmModel(NeuronLlamaModel):
    attention_cls = _GlmAttention
    graph_safe = False

    def make_rotary(self, config, device):
        rot = int(config.head_dim * _partial_rotary(config, 0.5))
        return RotaryEmbedding(rot, max(config.max_position_embeddings, config.neuron_config.seq_len), rope_theta_of(config),
                               rope_scaling_of(config), device=device)


class NeuronGlmForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronGlmModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        return fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)     # gate_up_proj ships fused ([gate; up])


# ---------------------------------------------------------------------------------------------------------- Cohere2
class NeuronCohere2Model(NeuronClassicModel):
    SPEC = dict(NeuronClassicModel.SPEC, norm_bias=False)

    def layer_spec(self, config, i):
        lt = getattr(config, "layer_types", None)
        if lt:
            sliding = lt[i] == "sliding_attention"
        else:
            sliding = bool((i + 1) % getattr(config, "sliding_window_pattern", 4))
        b = bool(getattr(config, "attention_bias", False))
        return dict(parallel=True, shared_norm=True, norm_bias=False, mlp="gated", act=getattr(config, "hidden_act", "silu"), qkv_bias=b,
                    o_bias=b, mlp_bias=False, rope_interleaved=True, use_rope=sliding,
                    sliding_window=getattr(config, "sliding_window", None) if sliding else None)


class NeuronCohere2ForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronCohere2Model

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        return fuse_qkv_and_gate_up(sd, config.num_hidden_layers)


# ---------------------------------------------------------------------------------------------------------- Apertus
class XIELU(nn.Module):
    """xIELU (arXiv 2411.13010): ``x>0: a_p x^2 + b x``, else ``a_n (expm1(min(x, eps)) - x) + b x`` with learned, softplus-
    parameterised a_p / a_n.  Elementwise, so it commutes with the column sharding of the up projection."""

    def __init__(self, dtype, device=None):
        super().__init__()
        self.alpha_p = nn.Parameter(torch.zeros(1, dtype=dtype, device=device), requires_grad=False)
        self.alpha_n = nn.Parameter(torch.zeros(1, dtype=dtype, device=device), requires_grad=False)
        self.register_buffer("beta", torch.tensor(0.5, dtype=dtype, device=device))
        self.register_buffer("eps", torch.tensor(-1e-6, dtype=dtype, device=device))

    def forward(self, x):
        a_p = F.softplus(self.alpha_p)
        a_n = self.beta + F.softplus(self.alpha_n)
        return torch.where(x > 0, a_p * x * x + self.beta * x, (torch.expm1(torch.min(x, self.eps)) - x) * a_n + self.beta * x)


class _XieluMLP(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        self.inner = PlainMLP(config.hidden_size, config.intermediate_size, "relu", dt, bias=getattr(config, "mlp_bias", False), device=device)
        self.act_fn = XIELU(dt, device)

    def forward(self, x, norm_weight=None, norm_eps=1e-6, norm_offset=0.0, residual=None, **kw):
        xn = ops.rmsnorm(x, norm_weight, norm_eps, norm_offset) if norm_weight is not None else x
        return self.inner.fc2(self.act_fn(self.inner.fc1(xn)), residual)


class _ApertusAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, qk_norm="rms_pre_rope", qk_norm_eps=config.rms_norm_eps, **over)


class NeuronApertusModel(NeuronLlamaModel):
    attention_cls = _ApertusAttention
    mlp_cls = _XieluMLP
    graph_safe = False


class NeuronApertusForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronApertusModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
        ren = ((".attention_layernorm.", ".input_layernorm."), (".feedforward_layernorm.", ".post_attention_layernorm."),
               (".self_attn.q_norm.", ".self_attn.q_layernorm."), (".self_attn.k_norm.", ".self_attn.k_layernorm."),
               (".mlp.up_proj.", ".mlp.inner.fc1."), (".mlp.down_proj.", ".mlp.inner.fc2."))
        out = {}
        for k, v in sd.items():
            for a, b in ren:
                k = k.replace(a, b)
            out[k] = v
        return out


# ---------------------------------------------------------------------------------------------------------- Nemotron
class NeuronNemotronModel(NeuronClassicModel):
    def make_rotary(self, config, device):
        rot = int(config.head_dim * _partial_rotary(config, 0.5))
        return RotaryEmbedding(rot, max(config.max_position_embeddings, config.neuron_config.seq_len), rope_theta_of(config),
                               rope_scaling_of(config), device=device)

    def layer_spec(self, config, i):
        ab, mb = bool(getattr(config, "attention_bias", False)), bool(getattr(config, "mlp_bias", False))
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "hidden_act", "relu2"), qkv_bias=ab, o_bias=ab, mlp_bias=mb)


class NeuronNemotronForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronNemotronModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
        sd = _rename_plain_mlp(sd, config.num_hidden_layers, "up_proj", "down_proj")
        for k in list(sd):       # LayerNorm1P: y = ln(x) * (1 + w) + b
            if k.endswith("layernorm.weight") or k == "norm.weight":
                sd[k] = sd[k] + 1.0
        return sd

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        sd["lm_head.weight"] = sd["embed_tokens.weight"].clone()


MORE_MODEL_TYPES = {"gemma": NeuronGemmaForCausalLM, "vaultgemma": NeuronVaultGemmaForCausalLM, "glm": NeuronGlmForCausalLM,
                    "cohere2": NeuronCohere2ForCausalLM, "apertus": NeuronApertusForCausalLM, "nemotron": NeuronNemotronForCausalLM}
