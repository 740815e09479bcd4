"""Llama-derivative families expressed as small deltas over ``NeuronLlamaModel``:

* **Phi-3 / Phi-3.5 / Phi-4-mini** — checkpoints already ship fused ``qkv_proj`` / ``gate_up_proj``; partial rotary.
* **Granite 3.x** — embedding / attention / residual / logits multipliers (folded into weights or scales at load).
* **SmolLM3** — NoPE on every ``no_rope_layer_interval``-th layer.
* **Seed-OSS** — q/k/v biases without an output bias, explicit head_dim.
* **OLMo-2 / OLMo-3** — post-norm blocks (``x + norm(f(x))``) and q/k RMSNorm over the whole projection.
* **Gemma-2** — Gemma-3 block without q/k norm, with attention / final logit soft-capping.
* **GLM-4 (0414)** — sandwich norms, partial interleaved rotary, fused gate_up in the checkpoint.
reference ports: contrib/models/{Phi-3-mini-4k-instruct, Phi-3.5-mini-instruct, granite-3.1-8b-instruct, SmolLM3-3B,
Seed-OSS-36B-Instruct, OLMo-2-1124-7B, OLMo-3-7B-Think, gemma-2-9b, GLM-4-9B-0414}/src."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops
from ...models.gemma3.modeling_gemma3 import Gemma3DecoderLayer, Gemma3InferenceConfig, NeuronGemma3ForCausalLM, NeuronGemma3Model
from ...models.llama.modeling_llama import (LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaMLP,
                                            NeuronLlamaModel)
from ...models.model_base import DecoderLayer
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from ...modules.attention import AttentionBase
from ...modules.mlp import GatedMLP
from ...modules.norm import RMSNorm
from ...modules.rope import RotaryEmbedding


# ---------------------------------------------------------------------------------------------------------- Phi-3
class NeuronPhi3ForCausalLM(NeuronLlamaForCausalLM):
    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        # qkv_proj is already [q; k; v], gate_up_proj already [gate; up]
        return {k: v for k, v in sd.items() if "rotary_emb.inv_freq" not in k}


# ---------------------------------------------------------------------------------------------------------- Granite
class NeuronGraniteModel(NeuronLlamaModel):
    def init_model(self, config):
        super().init_model(config)
        self.embed_scale = float(getattr(config, "embedding_multiplier", 1.0))

    def make_layer(self, config, i, rotary, device):
        nc = config.neuron_config
        attn = self.attention_cls(config, i, rotary, device=device, softmax_scale=float(getattr(config, "attention_multiplier", None)
                                                                                        or 1.0 / math.sqrt(config.head_dim)))
        return DecoderLayer(attn, self.mlp_cls(config, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device), i)


class NeuronGraniteForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronGraniteModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        """``h + residual_multiplier * f(h)`` and ``logits / logits_scaling`` are linear in the last projection of f / in the
        lm_head: fold them into those weights so the fused (+residual) kernels apply unchanged."""
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers)
        rm = float(getattr(config, "residual_multiplier", 1.0))
        for i in range(config.num_hidden_layers):
            for k in (f"layers.{i}.self_attn.o_proj.weight", f"layers.{i}.mlp.down_proj.weight", f"layers.{i}.self_attn.o_proj.bias",
                      f"layers.{i}.mlp.down_proj.bias"):
                if k in sd:
                    sd[k] = (sd[k].float() * rm).to(sd[k].dtype)
        ls = float(getattr(config, "logits_scaling", 1.0))
        if "lm_head.weight" not in sd and "embed_tokens.weight" in sd:
            sd["lm_head.weight"] = sd["embed_tokens.weight"].clone()
        sd["lm_head.weight"] = (sd["lm_head.weight"].float() / ls).to(sd["lm_head.weight"].dtype)
        return sd

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        pass   # handled (with the logits scaling) in the conversion


# ---------------------------------------------------------------------------------------------------------- SmolLM3
class NeuronSmolLM3Model(NeuronLlamaModel):
    def make_layer(self, config, i, rotary, device):
        nc = config.neuron_config
        nrl = getattr(config, "no_rope_layers", None)
        if not nrl:
            step = getattr(config, "no_rope_layer_interval", 4)
            nrl = [int((j + 1) % step != 0) for j in range(config.num_hidden_layers)]
        attn = self.attention_cls(config, i, rotary, device=device, use_rope=bool(nrl[i]))
        return DecoderLayer(attn, self.mlp_cls(config, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, nc.torch_dtype, device=device), i)


class NeuronSmolLM3ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronSmolLM3Model


# ---------------------------------------------------------------------------------------------------------- Seed-OSS
class NeuronSeedOssAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, qkv_bias=bool(getattr(config, "attention_bias", True)),
                         o_bias=bool(getattr(config, "attention_out_bias", False)), **over)


class NeuronSeedOssModel(NeuronLlamaModel):
    attention_cls = NeuronSeedOssAttention


class NeuronSeedOssForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronSeedOssModel


# ---------------------------------------------------------------------------------------------------------- OLMo-2 / OLMo-3
class Olmo2Attention(AttentionBase):
    """q/k RMSNorm spans ALL heads of the projection (weight [n_heads * D]); with TP each rank normalises its own head slice with
    the matching weight slice but the statistics need the full vector -> one tiny all-reduce of the sum of squares."""

    def __init__(self, config, layer_idx, rotary_emb, device=None, sliding_window=None):
        super().__init__(config, hidden_size=config.hidden_size, num_attention_heads=config.num_attention_heads,
                         num_key_value_heads=config.num_key_value_heads, head_dim=config.head_dim, rotary_emb=rotary_emb,
                         qkv_bias=getattr(config, "attention_bias", False), o_bias=getattr(config, "attention_bias", False),
                         sliding_window=sliding_window, layer_idx=layer_idx, rms_norm_eps=config.rms_norm_eps, device=device)
        dt = config.neuron_config.torch_dtype
        plan, D = self.qkv_proj.plan, self.head_dim
        self.q_norm = nn.Parameter(torch.ones(self.n_q * D, dtype=dt, device=device), requires_grad=False)
        self.k_norm = nn.Parameter(torch.ones(self.n_kv * D, dtype=dt, device=device), requires_grad=False)
        from ...modules.gqa import _gather_heads
        self.q_norm.shard_fn = lambda full, rank: _gather_heads(full, plan.q_idx[rank], D, 0)
        self.k_norm.shard_fn = lambda full, rank: _gather_heads(full, plan.kv_idx[rank], D, 0)
        for p in (self.q_norm, self.k_norm):
            p.partition_dim, p.tp_group = 0, self.tp_group
        self.full_q, self.full_k = config.num_attention_heads * D, config.num_key_value_heads * D
        self.eps = config.rms_norm_eps

    def _simple(self):
        return False

    def _full_rms(self, x, w, full_width, B, T, n_heads):
        xf = x.reshape(B, T, -1).float()
        ss = xf.pow(2).sum(-1, keepdim=True)
        if self.tp_group.size > 1:
            from ...parallel import mappings
            # replicated kv heads are counted once per owner: scale by the replication factor
            rep = (n_heads * self.head_dim * self.tp_group.size) / full_width
            ss = mappings.all_reduce(ss, self.tp_group) / rep
        y = xf * torch.rsqrt(ss / full_width + self.eps) * w.float()
        return y.to(x.dtype).view(B, T, n_heads, self.head_dim)

    def _split_norm_rope(self, qkv, B, T, cos, sin, meta=None):
        D, nq, nkv = self.head_dim, self.n_q, self.n_kv
        q, k, v = qkv.reshape(B, T, nq + 2 * nkv, D).split([nq, nkv, nkv], dim=2)
        q = self._full_rms(q, self.q_norm, self.full_q, B, T, nq)
        k = self._full_rms(k, self.k_norm, self.full_k, B, T, nkv)
        if cos is not None:
            q, k = ops.apply_rope(q, cos, sin, False), ops.apply_rope(k, cos, sin, False)
        return q, k, v


class Olmo2DecoderLayer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, rotary, device=None, sliding_window=None):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        self.self_attn = Olmo2Attention(config, i, rotary, device, sliding_window)
        self.mlp = GatedMLP(config.hidden_size, config.intermediate_size, config.hidden_act, dt, device=device)
        self.post_attention_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.post_feedforward_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.layer_idx = i

    def forward(self, h, meta, kv_mgr, lora=None):
        h = h + self.post_attention_layernorm(self.self_attn(h, meta, kv_mgr))
        return h + self.post_feedforward_layernorm(self.mlp(h))


class NeuronOlmo2Model(NeuronLlamaModel):
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        lt = getattr(config, "layer_types", None)
        sw = getattr(config, "sliding_window", None) if (lt and lt[i] == "sliding_attention") else None
        return Olmo2DecoderLayer(config, i, rotary, device, sw)


class NeuronOlmo2ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronOlmo2Model

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers)
        return {k.replace(".self_attn.q_norm.weight", ".self_attn.q_norm").replace(".self_attn.k_norm.weight", ".self_attn.k_norm"): v
                for k, v in sd.items()}


NeuronOlmo3ForCausalLM = NeuronOlmo2ForCausalLM


# ---------------------------------------------------------------------------------------------------------- Gemma-2
class Gemma2DecoderLayer(Gemma3DecoderLayer):
    def __init__(self, config, i, rope_local, rope_global, device=None):
        super().__init__(config, i, rope_local, rope_global, device)
        a = self.self_attn
        a.qk_norm = None                       # Gemma-2 has no q/k norm
        del a.q_layernorm, a.k_layernorm


class NeuronGemma2Model(NeuronGemma3Model):
    def init_model(self, config):
        if not getattr(config, "layer_types", None):   # Gemma-2: even layers slide
            config.layer_types = ["sliding_attention" if (i % 2 == 0) else "full_attention" for i in range(config.num_hidden_layers)]
        super().init_model(config)
        nc = config.neuron_config
        maxpos = max(config.max_position_embeddings, nc.seq_len)
        from ...models.llama.modeling_llama import rope_theta_of
        rope = RotaryEmbedding(config.head_dim, maxpos, rope_theta_of(config, 10000.0), None, device=self.device_)
        self.layers = nn.ModuleList([Gemma2DecoderLayer(config, i, rope, rope, self.device_) for i in range(config.num_hidden_layers)])


class NeuronGemma2ForCausalLM(NeuronGemma3ForCausalLM):
    _model_cls = NeuronGemma2Model


# ---------------------------------------------------------------------------------------------------------- GLM-4 (0414)
class Glm4DecoderLayer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        nc = config.neuron_config
        dt = nc.torch_dtype
        self.self_attn = AttentionBase(config, hidden_size=config.hidden_size, num_attention_heads=config.num_attention_heads,
                                       num_key_value_heads=config.num_key_value_heads, head_dim=config.head_dim, rotary_emb=rotary,
                                       qkv_bias=getattr(config, "attention_bias", True), o_bias=False, rope_interleaved=True,
                                       layer_idx=i, rms_norm_eps=config.rms_norm_eps, device=device)
        self.mlp = GatedMLP(config.hidden_size, config.intermediate_size, config.hidden_act, dt, device=device)
        mk = lambda: RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)   # noqa: E731
        self.input_layernorm, self.post_attention_layernorm = mk(), mk()
        self.post_self_attn_layernorm, self.post_mlp_layernorm = mk(), mk()
        self.layer_idx = i

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.input_layernorm
        a = self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon)
        h = h + self.post_self_attn_layernorm(a)
        n = self.post_attention_layernorm
        m = self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon)
        return h + self.post_mlp_layernorm(m)


class NeuronGlm4Model(NeuronLlamaModel):
    graph_safe = False

    def make_rotary(self, config, device):
        rot = int(config.head_dim * getattr(config, "partial_rotary_factor", 0.5))
        from ...models.llama.modeling_llama import rope_scaling_of, rope_theta_of
        return RotaryEmbedding(rot, max(config.max_position_embeddings, config.neuron_config.seq_len), rope_theta_of(config),
                               rope_scaling_of(config), device=device)

    def make_layer(self, config, i, rotary, device):
        return Glm4DecoderLayer(config, i, rotary, device)


class NeuronGlm4ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronGlm4Model

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)    # gate_up_proj ships fused
        return sd


# ---------------------------------------------------------------------------------------------------------- Helium / ERNIE-4.5
class _InterleavedRopeAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, rope_interleaved=True, **over)


class NeuronHeliumModel(NeuronLlamaModel):
    """Llama block with GPT-J style (pairwise / interleaved) rotary."""
    attention_cls = _InterleavedRopeAttention
    graph_safe = False


class NeuronHeliumForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronHeliumModel


NeuronErnie4_5ForCausalLM = NeuronHeliumForCausalLM


# ---------------------------------------------------------------------------------------------------------- Arcee AFM
class _Relu2MLP(nn.Module):
    """Non-gated ``down(relu(up(x))^2)`` (AFM-4.5B); keeps the GatedMLP calling convention (fused norm, residual)."""

    def __init__(self, config, device=None):
        super().__init__()
        from ...modules.mlp import PlainMLP
        self.inner = PlainMLP(config.hidden_size, config.intermediate_size, "relu2", config.neuron_config.torch_dtype,
                              bias=getattr(config, "mlp_bias", False), device=device)

    def forward(self, x, norm_weight=None, norm_eps=1e-6, norm_offset=0.0, residual=None, **kw):
        xn = ops.rmsnorm(x, norm_weight, norm_eps, norm_offset) if norm_weight is not None else x
        return self.inner(xn, residual)


class NeuronArceeModel(NeuronLlamaModel):
    mlp_cls = _Relu2MLP
    graph_safe = False


class NeuronArceeForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronArceeModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
        return {k.replace(".mlp.up_proj.", ".mlp.inner.fc1.").replace(".mlp.down_proj.", ".mlp.inner.fc2."): v for k, v in sd.items()}



CONTRIB_MODEL_TYPES = {
    "helium": NeuronHeliumForCausalLM, "ernie4_5": NeuronErnie4_5ForCausalLM, "arcee": NeuronArceeForCausalLM,
    "phi3": NeuronPhi3ForCausalLM, "granite": NeuronGraniteForCausalLM, "smollm3": NeuronSmolLM3ForCausalLM,
    "seed_oss": NeuronSeedOssForCausalLM, "olmo2": NeuronOlmo2ForCausalLM, "olmo3": NeuronOlmo3ForCausalLM,
    "gemma2": NeuronGemma2ForCausalLM, "glm4": NeuronGlm4ForCausalLM,
}
