"""ALiBi decoders (no position embeddings; every head adds ``slope_h * (key_pos - query_pos)`` to its attention logits):

* **BLOOM** — LayerNorm on the embeddings, per-head interleaved fused QKV, biases everywhere, tanh-GELU MLP.
* **MPT** — bias-free LayerNorms and projections, ``[q; k; v]`` fused ``Wqkv``, GELU MLP, ``alibi_bias_max`` slopes.

ALiBi runs on a masked fp32 attention path written here (`AlibiAttention`) over the engine's contiguous KV cache: the bias is a rank-1
term that the flash kernels of `csrc/attention.cu` do not take yet (adding a per-head slope argument to them is a small change; these
families are not on a benchmark path).  Checked against Hugging Face in tests/test_contrib_cpu.py."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...modules.attention import AttentionBase
from .classic_family import ClassicInferenceConfig, NeuronClassicModel, _ClassicCausalLM


def alibi_slopes(n_heads: int, bias_max: float = 8.0, mpt_order: bool = False) -> torch.Tensor:
    """Geometric slopes ``2^(-bias_max * i / n)``; for head counts that are not a power of two BLOOM appends every other slope of the
    next power of two, MPT interleaves them."""
    p2 = 2 ** math.floor(math.log2(n_heads))
    if mpt_order:
        p2 = 2 ** math.ceil(math.log2(n_heads))
        s = torch.pow(2.0, -torch.arange(1, p2 + 1, dtype=torch.float32) * (bias_max / p2))
        return s if p2 == n_heads else torch.cat([s[1::2], s[::2]])[:n_heads]
    s = torch.pow(2.0, -torch.arange(1, p2 + 1, dtype=torch.float32) * (bias_max / p2))
    if p2 != n_heads:
        extra = torch.pow(2.0, -torch.arange(1, 2 * (n_heads - p2), 2, dtype=torch.float32) * (bias_max / (2 * p2)))
        s = torch.cat([s, extra])
    return s


class AlibiAttention(AttentionBase):
    def __init__(self, config, *, alibi_mpt: bool = False, **kw):
        super().__init__(config, **kw)
        full = alibi_slopes(config.num_attention_heads, float(getattr(config, "alibi_bias_max", 8.0)), alibi_mpt)
        r = self.tp_group.rank
        self.register_buffer("slopes", full[r * self.n_q:(r + 1) * self.n_q].clone().to(kw.get("device")), persistent=False)

    def forward(self, hidden, meta, kv_mgr, norm_weight=None, norm_eps=None, norm_offset: float = 0.0, residual=None, lora=None):
        if meta.slot_mapping is not None or meta.has_prefix or meta.active_mask is not None or self.neuron_config.padding_side != "right":
            raise NotImplementedError("ALiBi attention: contiguous KV cache, right padding, one token per decode step")
        B, T, _ = hidden.shape
        qkv = self.qkv_proj(hidden, norm_weight, norm_eps if norm_eps is not None else self.rms_norm_eps, norm_offset)
        q, k, v = self._split_norm_rope(qkv, B, T, None, None, meta)                       # [B, T, heads, D]
        if meta.lines is None:
            meta.lines = kv_mgr.lines_for(meta.seq_ids)
        kv_mgr.update(self.layer_idx, k, v, meta.seq_ids, meta.write_positions, meta.lines)
        if meta.is_prefill:
            kk, vv = k.transpose(1, 2), v.transpose(1, 2)
            qpos = torch.arange(T, device=q.device).view(1, T).expand(B, T)
        else:
            k_cache, v_cache = kv_mgr.get_kv_by_layer_id(self.layer_idx)
            kk, vv = k_cache[meta.lines.long()], v_cache[meta.lines.long()]               # [B, Hkv, S, D]
            qpos = meta.position_ids.long()
        S = kk.shape[2]
        kpos = torch.arange(S, device=q.device)
        rel = (kpos.view(1, 1, S) - qpos.view(B, T, 1)).float()                            # <= 0 on the visible keys
        rep = self.n_q // self.n_kv
        kk, vv = kk.float().repeat_interleave(rep, 1), vv.float().repeat_interleave(rep, 1)
        s = torch.matmul(q.transpose(1, 2).float(), kk.transpose(-1, -2)) * self.scale + self.slopes.view(1, -1, 1, 1) * rel.unsqueeze(1)
        s = s.masked_fill((rel > 0).unsqueeze(1), float("-inf"))
        o = torch.matmul(torch.softmax(s, -1), vv).transpose(1, 2).to(hidden.dtype)        # [B, T, heads, D]
        return self._finish(o.reshape(B, T, self.n_q * self.head_dim), residual, lora, meta)


class _AlibiModel(NeuronClassicModel):
    graph_safe = False
    alibi_mpt = False

    def make_rotary(self, config, device):
        return None

    def layer_spec(self, config, i):
        mpt = self.alibi_mpt
        return dict(self.SPEC, attn_cls=lambda cfg, **kw: AlibiAttention(cfg, alibi_mpt=mpt, **kw))


# ---- BLOOM -----------------------------------------------------------------------------------------------------------------------------
class BloomInferenceConfig(ClassicInferenceConfig):
    attribute_map = {"n_head": "num_attention_heads", "n_layer": "num_hidden_layers", "layer_norm_epsilon": "rms_norm_eps"}

    def add_derived_config(self):
        if getattr(self, "apply_residual_connection_post_layernorm", False):
            raise NotImplementedError("BLOOM with apply_residual_connection_post_layernorm")
        self.max_position_embeddings = getattr(self, "max_position_embeddings", None) or getattr(self, "seq_length", 2048)
        super().add_derived_config()


class NeuronBloomModel(_AlibiModel):
    SPEC = dict(parallel=False, norm_bias=True, mlp="plain", act="gelu_pytorch_tanh", qkv_bias=True, o_bias=True, mlp_bias=True)

    def init_model(self, config):
        super().init_model(config)
        self.embed_layernorm = nn.LayerNorm(config.hidden_size, eps=config.rms_norm_eps, dtype=config.neuron_config.torch_dtype, device=self.device_)
        for p in self.embed_layernorm.parameters():
            p.requires_grad_(False)

    def embed(self, input_ids, inputs_embeds=None, vision_embeddings=None, vision_mask=None):
        return self.embed_layernorm(super().embed(input_ids, inputs_embeds, vision_embeddings, vision_mask))


class NeuronBloomForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronBloomModel
    _STATE_DICT_MODEL_PREFIX = "transformer."

    @classmethod
    def get_config_cls(cls):
        return BloomInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        nh, D = config.num_attention_heads, config.hidden_size // config.num_attention_heads
        ren = (("word_embeddings_layernorm.", "embed_layernorm."), ("word_embeddings.", "embed_tokens."), ("ln_f.", "norm."), ("h.", "layers."),
               (".self_attention.dense.", ".self_attn.o_proj."), (".mlp.dense_h_to_4h.", ".mlp.fc1."), (".mlp.dense_4h_to_h.", ".mlp.fc2."))
        out = {}
        for k, v in sd.items():
            for a, b in ren:
                if k.startswith(a) or a.startswith(".") and a in k:
                    k = k.replace(a, b, 1)
            if ".self_attention.query_key_value." in k:                                   # rows are [head, (q, k, v), D]
                w = v.view(nh, 3, D, *v.shape[1:])
                v = torch.cat([w[:, j].reshape(nh * D, *v.shape[1:]) for j in range(3)], 0)
                k = k.replace(".self_attention.query_key_value.", ".self_attn.qkv_proj.")
            out[k] = v
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


# ---- MPT -------------------------------------------------------------------------------------------------------------------------------
class MptInferenceConfig(ClassicInferenceConfig):
    attribute_map = {"d_model": "hidden_size", "n_heads": "num_attention_heads", "n_layers": "num_hidden_layers",
                     "max_seq_len": "max_position_embeddings", "layer_norm_epsilon": "rms_norm_eps"}

    def add_derived_config(self):
        ac = getattr(self, "attn_config", None) or {}
        ac = ac if isinstance(ac, dict) else ac.to_dict()
        if not ac.get("alibi", True) or ac.get("qk_ln", False) or ac.get("clip_qkv") or not getattr(self, "no_bias", True):
            raise NotImplementedError("MPT variants without ALiBi / with q-k LayerNorm, clip_qkv or biases")
        self.alibi_bias_max = float(ac.get("alibi_bias_max", 8))
        self.intermediate_size = int(getattr(self, "expansion_ratio", 4) * self.hidden_size)
        super().add_derived_config()


class NeuronMptModel(_AlibiModel):
    SPEC = dict(parallel=False, norm_bias=False, mlp="plain", act="gelu", qkv_bias=False, o_bias=False, mlp_bias=False)
    alibi_mpt = True


class NeuronMptForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronMptModel
    _STATE_DICT_MODEL_PREFIX = "transformer."

    @classmethod
    def get_config_cls(cls):
        return MptInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        ren = (("wte.", "embed_tokens."), ("norm_f.", "norm."), ("blocks.", "layers."), (".norm_1.", ".input_layernorm."),
               (".norm_2.", ".post_attention_layernorm."), (".attn.Wqkv.", ".self_attn.qkv_proj."), (".attn.out_proj.", ".self_attn.o_proj."),
               (".ffn.up_proj.", ".mlp.fc1."), (".ffn.down_proj.", ".mlp.fc2."))
        out = {}
        for k, v in sd.items():
            for a, b in ren:
                if k.startswith(a) or a.startswith(".") and a in k:
                    k = k.replace(a, b, 1)
            out[k] = v
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


ALIBI_MODEL_TYPES = {"bloom": NeuronBloomForCausalLM, "mpt": NeuronMptForCausalLM}
