"""More community-hub decoders, each a small delta over the Llama / classic blocks:

* **Gemma (v1)** — offset RMSNorm (1 + w), sqrt(H)-scaled embeddings, GeGLU, explicit head_dim, tied head.
* **VaultGemma** — Gemma-2 attention (soft-cap, ``query_pre_attn_scalar``, sliding / global layers) in a plain pre-norm block.
* **GLM-4-9B-chat (``glm``)** — partial (50 %) interleaved rotary, q/k/v biases, fused gate_up in the checkpoint.
* **Cohere2 / Command-R7B** — Cohere block; sliding-window layers use RoPE, full-attention layers have NO position encoding.
* **Apertus** — per-head q/k RMSNorm, non-gated MLP with the learned xIELU activation.
* **Persimmon** — per-head LayerNorm (with bias) on q and k, head-interleaved fused QKV, partial rotary, squared-ReLU MLP.
* **XGLM** — fairseq-style sinusoidal position table (offset 2), sqrt(H)-scaled embeddings, pre-LN GELU block.
* **CodeGen** — GPT-J block whose fused QKV is laid out in 4 "logical core" groups of [q | v | k].
* **OpenAI GPT (GPT-1)** — post-LayerNorm blocks, learned positions, Conv1D weights, no final norm, tied head.
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
from ...modules.attention import AttentionBase
from ...modules.mlp import PlainMLP
from ...modules.norm import RMSNorm
from ...modules.rope import RotaryEmbedding
from .classic_family import (ClassicInferenceConfig, GPTJInferenceConfig, NeuronClassicModel, NeuronGPTJForCausalLM, NeuronGPTJModel,
                             _ClassicCausalLM, _rename_plain_mlp)


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


# ---------------------------------------------------------------------------------------------------------- Persimmon
class _PersimmonAttention(AttentionBase):
    def __init__(self, config, **kw):
        super().__init__(config, **kw)
        dt, eps = config.neuron_config.torch_dtype, config.rms_norm_eps
        self.q_layernorm = nn.LayerNorm(self.head_dim, eps=eps, dtype=dt, device=kw.get("device"))
        self.k_layernorm = nn.LayerNorm(self.head_dim, eps=eps, dtype=dt, device=kw.get("device"))
        self.per_head_ln = bool(getattr(config, "qk_layernorm", True))
        for p in (*self.q_layernorm.parameters(), *self.k_layernorm.parameters()):
            p.requires_grad_(False)

    def _simple(self):
        return False

    def _split_norm_rope(self, qkv, B, T, cos, sin, meta=None):
        D, nq, nkv = self.head_dim, self.n_q, self.n_kv
        q, k, v = qkv.reshape(B, T, nq + 2 * nkv, D).split([nq, nkv, nkv], dim=2)
        if self.per_head_ln:
            q, k = self.q_layernorm(q), self.k_layernorm(k)
        if cos is not None:
            q, k = ops.apply_rope(q, cos, sin, False), ops.apply_rope(k, cos, sin, False)
        return q, k, v


class NeuronPersimmonModel(NeuronClassicModel):
    def layer_spec(self, config, i):
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "hidden_act", "relu2"), qkv_bias=True, o_bias=True,
                    mlp_bias=True, attn_cls=_PersimmonAttention)


class NeuronPersimmonForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronPersimmonModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        nh, D = config.num_attention_heads, config.hidden_size // config.num_attention_heads
        out = {}
        for k, v in sd.items():
            if ".self_attn.query_key_value." in k:        # [heads, (q,k,v), D, ...] -> [q heads; k heads; v heads]
                w = v.view(nh, 3, D, *v.shape[1:])
                v = torch.cat([w[:, j].reshape(nh * D, *v.shape[1:]) for j in range(3)], 0)
                k = k.replace(".query_key_value.", ".qkv_proj.")
            k = (k.replace(".self_attn.dense.", ".self_attn.o_proj.").replace(".mlp.dense_h_to_4h.", ".mlp.fc1.")
                 .replace(".mlp.dense_4h_to_h.", ".mlp.fc2.").replace("final_layernorm.", "norm."))
            out[k] = v
        return out


# ---------------------------------------------------------------------------------------------------------- XGLM
class XGLMInferenceConfig(ClassicInferenceConfig):
    attribute_map = {"d_model": "hidden_size", "attention_heads": "num_attention_heads", "num_layers": "num_hidden_layers",
                     "ffn_dim": "intermediate_size"}


def _fairseq_sinusoids(n, dim, padding_idx=None):
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000.0) / (half - 1)))
    ang = torch.arange(n, dtype=torch.float32).unsqueeze(1) * freq.unsqueeze(0)
    tab = torch.cat([ang.sin(), ang.cos()], 1)
    if dim % 2:
        tab = torch.cat([tab, torch.zeros(n, 1)], 1)
    if padding_idx is not None:
        tab[padding_idx] = 0
    return tab


class NeuronXGLMModel(NeuronClassicModel):
    learned_positions = True          # a fixed table, but looked up the same way
    position_offset = 2

    def init_model(self, config):
        super().init_model(config)
        self.embed_scale = float(config.hidden_size ** 0.5) if getattr(config, "scale_embedding", True) else 1.0

    def embed(self, input_ids, inputs_embeds=None, vision_embeddings=None, vision_mask=None):
        h = self.embed_tokens(input_ids) * self.embed_scale
        pos = self._pos if self._pos is not None else torch.arange(input_ids.shape[1], device=input_ids.device).unsqueeze(0)
        return h + self.embed_positions((pos.long() + self.position_offset).clamp(0, self.embed_positions.num_embeddings - 1))

    def layer_spec(self, config, i):
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "activation_function", "gelu"), qkv_bias=True, o_bias=True,
                    mlp_bias=True)


class NeuronXGLMForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronXGLMModel

    @classmethod
    def get_config_cls(cls):
        return XGLMInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k.startswith("layers."):
                k = (k.replace(".self_attn.out_proj.", ".self_attn.o_proj.").replace(".self_attn_layer_norm.", ".input_layernorm.")
                     .replace(".final_layer_norm.", ".post_attention_layernorm.").replace(".fc1.", ".mlp.fc1.").replace(".fc2.", ".mlp.fc2."))
            elif k.startswith("layer_norm."):
                k = k.replace("layer_norm.", "norm.")
            elif k.startswith("embed_positions."):
                continue
            out[k] = v
        out = fuse_qkv_and_gate_up(out, config.num_hidden_layers, fuse_mlp=False)
        tab = _fairseq_sinusoids(config.max_position_embeddings + 2, config.hidden_size, getattr(config, "pad_token_id", 1))
        out["embed_positions.weight"] = tab.to(out["embed_tokens.weight"].dtype)
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


# ---------------------------------------------------------------------------------------------------------- CodeGen
class NeuronCodeGenForCausalLM(NeuronGPTJForCausalLM):
    _model_cls = NeuronGPTJModel

    @classmethod
    def get_config_cls(cls):
        return GPTJInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        mp = 4
        out = {}
        for k, v in sd.items():
            if k.endswith(".attn.qkv_proj.weight"):
                H = v.shape[1]
                w = v.view(mp, 3, H // mp, H)                  # per logical core: [q | v | k]
                q, vv, kk = (w[:, j].reshape(H, H) for j in range(3))
                v = torch.cat([q, kk, vv], 0)
                k = k.replace(".attn.qkv_proj.", ".self_attn.qkv_proj.")
            out[k] = v
        return NeuronGPTJForCausalLM.convert_hf_to_neuron_state_dict(out, config)


# ---------------------------------------------------------------------------------------------------------- OpenAI GPT (GPT-1)
class NeuronOpenAIGPTModel(NeuronClassicModel):
    learned_positions = True

    def layer_spec(self, config, i):
        return dict(parallel=False, post_ln=True, norm_bias=True, mlp="plain", act=getattr(config, "afn", "gelu_new"), qkv_bias=True, o_bias=True,
                    mlp_bias=True)

    def init_model(self, config):
        super().init_model(config)
        self.norm = nn.Identity()


class NeuronOpenAIGPTForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronOpenAIGPTModel
    _STATE_DICT_MODEL_PREFIX = "transformer."

    @classmethod
    def get_config_cls(cls):
        return GPTJInferenceConfig          # n_embd / n_head / n_layer / n_positions aliases

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"):
                continue
            if any(k.endswith(s) for s in ("c_attn.weight", "attn.c_proj.weight", "c_fc.weight", "mlp.c_proj.weight")):
                v = v.t().contiguous()                      # Conv1D stores [in, out]
            k = k.replace("h.", "layers.", 1) if k.startswith("h.") else k
            k = (k.replace(".attn.c_attn.", ".self_attn.qkv_proj.").replace(".attn.c_proj.", ".self_attn.o_proj.").replace(".mlp.c_fc.", ".mlp.fc1.")
                 .replace(".mlp.c_proj.", ".mlp.fc2.").replace(".ln_1.", ".input_layernorm.").replace(".ln_2.", ".post_attention_layernorm."))
            out[k.replace("tokens_embed.", "embed_tokens.").replace("positions_embed.", "embed_positions.")] = v
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


MORE_MODEL_TYPES = {"openai-gpt": NeuronOpenAIGPTForCausalLM,
                    "persimmon": NeuronPersimmonForCausalLM, "xglm": NeuronXGLMForCausalLM, "codegen": NeuronCodeGenForCausalLM,
                    "gemma": NeuronGemmaForCausalLM, "vaultgemma": NeuronVaultGemmaForCausalLM, "glm": NeuronGlmForCausalLM,
                    "cohere2": NeuronCohere2ForCausalLM, "apertus": NeuronApertusForCausalLM, "nemotron": NeuronNemotronForCausalLM}
