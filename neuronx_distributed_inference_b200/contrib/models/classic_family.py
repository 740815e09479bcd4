"""Pre-Llama style decoders on one configurable block: LayerNorm (with or without bias), sequential or *parallel* residual,
plain (fc1-act-fc2) or gated MLP, learned positions or (partial / interleaved) rotary.

* **StarCoder2** — sequential, LayerNorm+bias, plain GELU-tanh MLP with biases, GQA, optional sliding window.
* **StableLM-2** — sequential, LayerNorm, gated SiLU MLP, partial rotary (25 %).
* **Cohere Command-R** — parallel block fed by ONE bias-free LayerNorm, interleaved rotary, logit scale, tied embeddings.
* **GPT-NeoX / Pythia** — parallel residual with two LayerNorms, per-head interleaved fused QKV, partial rotary, plain GELU MLP.
* **GPT-2** — learned absolute positions, Conv1D ([in,out]) weights, tied head.
reference ports: contrib/models/{starcoder2-3b, stablelm-2-1_6b, c4ai-command-r7b-12-2024, pythia-2.8b, gpt2}/src."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...models.application_base import NeuronBaseForCausalLM
from ...models.llama.modeling_llama import LlamaInferenceConfig, rope_scaling_of, rope_theta_of
from ...models.model_base import NeuronBaseModel
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from ...modules.attention import AttentionBase
from ...modules.mlp import GatedMLP, PlainMLP
from ...modules.rope import RotaryEmbedding
from ...parallel.layers import ColumnParallelLinear, ParallelEmbedding


class ClassicInferenceConfig(LlamaInferenceConfig):
    def get_required_attributes(self):
        return ["hidden_size", "num_attention_heads", "num_hidden_layers", "vocab_size"]

    def add_derived_config(self):
        if not hasattr(self, "num_key_value_heads") or self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        if not hasattr(self, "intermediate_size") or self.intermediate_size is None:
            self.intermediate_size = 4 * self.hidden_size
        if not hasattr(self, "rms_norm_eps"):
            self.rms_norm_eps = getattr(self, "layer_norm_eps", getattr(self, "norm_epsilon", getattr(self, "layer_norm_epsilon", 1e-5)))
        if not hasattr(self, "max_position_embeddings"):
            self.max_position_embeddings = getattr(self, "n_positions", 2048)
        super().add_derived_config()


class ClassicDecoderLayer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, rotary, spec, device=None):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        H, eps = config.hidden_size, config.rms_norm_eps
        self.parallel, self.shared_norm, self.post_ln = spec["parallel"], spec.get("shared_norm", False), spec.get("post_ln", False)
        self.self_attn = spec.get("attn_cls", AttentionBase)(config, hidden_size=H, num_attention_heads=config.num_attention_heads,
                                       num_key_value_heads=config.num_key_value_heads, head_dim=config.head_dim, rotary_emb=rotary,
                                       qkv_bias=spec["qkv_bias"], o_bias=spec["o_bias"], use_rope=rotary is not None and spec.get("use_rope", True),
                                       rope_interleaved=spec.get("rope_interleaved", False),
                                       sliding_window=spec.get("sliding_window"), softmax_scale=spec.get("softmax_scale"),
                                       clip_qkv=spec.get("clip_qkv"), layer_idx=i, device=device)
        if callable(spec["mlp"]):                     # MoE (or any custom) feed-forward: factory(config, device)
            self.mlp = spec["mlp"](config, device)
            self.mlp_is_moe = True
        elif spec["mlp"] == "gated":
            self.mlp = GatedMLP(H, config.intermediate_size, spec["act"], dt, bias=spec["mlp_bias"], device=device)
        else:
            self.mlp = PlainMLP(H, config.intermediate_size, spec["act"], dt, bias=spec["mlp_bias"], device=device)
        self.input_layernorm = nn.LayerNorm(H, eps=eps, bias=spec["norm_bias"], dtype=dt, device=device)
        self.post_attention_layernorm = None if self.shared_norm else nn.LayerNorm(H, eps=eps, bias=spec["norm_bias"], dtype=dt, device=device)
        for p in self.parameters():
            p.requires_grad_(False)
        self.layer_idx = i

    def forward(self, h, meta, kv_mgr, lora=None):
        if self.post_ln:                                   # GPT-1: norms AFTER the residual sums
            h = self.input_layernorm(h + self.self_attn(h, meta, kv_mgr))
            return self.post_attention_layernorm(h + self.mlp(h))
        x1 = self.input_layernorm(h)
        a = self.self_attn(x1, meta, kv_mgr)
        if self.parallel:
            return h + a + self.mlp(x1 if self.shared_norm else self.post_attention_layernorm(h))
        h = h + a
        return h + self.mlp(self.post_attention_layernorm(h))


class NeuronClassicModel(NeuronBaseModel):
    """Subclasses set ``SPEC`` (block options) and may override ``layer_spec(config, i)``."""
    graph_safe = False
    SPEC = dict(parallel=False, norm_bias=True, mlp="plain", act="gelu", qkv_bias=True, o_bias=True, mlp_bias=True)
    learned_positions = False
    position_offset = 0

    def setup_attr_for_model(self, config):
        nc = config.neuron_config
        self.tp_degree, self.hidden_size = nc.tp_degree, config.hidden_size
        self.num_attention_heads, self.num_key_value_heads = config.num_attention_heads, config.num_key_value_heads
        self.max_batch_size, self.buckets = nc.max_batch_size, nc.buckets

    def layer_spec(self, config, i):
        return dict(self.SPEC)

    def make_rotary(self, config, device):
        rp = getattr(config, "rope_parameters", None) or {}
        frac = getattr(config, "partial_rotary_factor", None) or getattr(config, "rotary_pct", None) or (
            rp.get("partial_rotary_factor") if isinstance(rp, dict) else None) or 1.0
        rot = int(config.head_dim * float(frac))
        return RotaryEmbedding(rot, max(config.max_position_embeddings, config.neuron_config.seq_len), rope_theta_of(config),
                               rope_scaling_of(config), device=device)

    def init_model(self, config):
        nc, dev = config.neuron_config, self.device_
        dt = nc.torch_dtype
        self.embed_tokens = ParallelEmbedding(config.vocab_size, config.hidden_size, None, dtype=dt, device=dev,
                                              shard_across_embedding=True, pad=True, tensor_model_parallel_group=self.tp_group)
        rotary = None
        if self.learned_positions:
            self.embed_positions = nn.Embedding(config.max_position_embeddings + self.position_offset, config.hidden_size, dtype=dt, device=dev)
            self.embed_positions.weight.requires_grad_(False)
        else:
            rotary = self.make_rotary(config, dev)
        self.layers = nn.ModuleList([ClassicDecoderLayer(config, i, rotary, self.layer_spec(config, i), dev)
                                     for i in range(config.num_hidden_layers)])
        self.norm = nn.LayerNorm(config.hidden_size, eps=config.rms_norm_eps, bias=self.SPEC["norm_bias"], dtype=dt, device=dev)
        for p in self.norm.parameters():
            p.requires_grad_(False)
        self.lm_head = ColumnParallelLinear(config.hidden_size, config.vocab_size, bias=bool(getattr(self, "lm_head_bias", False)),
                                            gather_output=False, dtype=dt, device=dev, pad=True, tensor_model_parallel_group=self.tp_group)
        self.logit_scale = float(getattr(config, "logit_scale", 1.0) or 1.0)

    def forward(self, input_ids, attention_mask=None, position_ids=None, *a, **kw):
        self._pos = position_ids
        return super().forward(input_ids, attention_mask, position_ids, *a, **kw)

    def embed(self, input_ids, inputs_embeds=None, vision_embeddings=None, vision_mask=None):
        h = self.embed_tokens(input_ids)
        if self.learned_positions:
            pos = self._pos if self._pos is not None else torch.arange(input_ids.shape[1], device=input_ids.device).unsqueeze(0)
            h = h + self.embed_positions((pos.long() + self.position_offset).clamp(0, self.embed_positions.num_embeddings - 1))
        return h

    def final_hidden(self, h):
        return self.norm(h)

    def compute_logits(self, h):
        logits = self.lm_head(self.norm(h))
        return logits if self.logit_scale == 1.0 else logits * self.logit_scale


class _ClassicCausalLM(NeuronBaseForCausalLM):
    @classmethod
    def get_config_cls(cls):
        return ClassicInferenceConfig

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        sd["lm_head.weight"] = sd["embed_tokens.weight"].clone()


def _rename_plain_mlp(sd, n_layers, fc1, fc2):
    for i in range(n_layers):
        for suf in ("weight", "bias"):
            for src, dst in ((fc1, "fc1"), (fc2, "fc2")):
                k = f"layers.{i}.mlp.{src}.{suf}"
                if k in sd:
                    sd[f"layers.{i}.mlp.{dst}.{suf}"] = sd.pop(k)
    return sd


# ---- StarCoder2 ------------------------------------------------------------------------------------------------------
class NeuronStarcoder2Model(NeuronClassicModel):
    def layer_spec(self, config, i):
        b = bool(getattr(config, "use_bias", True))
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "hidden_act", "gelu_pytorch_tanh"), qkv_bias=b, o_bias=b,
                    mlp_bias=b, sliding_window=getattr(config, "sliding_window", None))


class NeuronStarcoder2ForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronStarcoder2Model

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
        return _rename_plain_mlp(sd, config.num_hidden_layers, "c_fc", "c_proj")


# ---- StableLM ---------------------------------------------------------------------------------------------------------
class NeuronStableLmModel(NeuronClassicModel):
    def layer_spec(self, config, i):
        return dict(parallel=bool(getattr(config, "use_parallel_residual", False)), norm_bias=True, mlp="gated",
                    act=getattr(config, "hidden_act", "silu"), qkv_bias=bool(getattr(config, "use_qkv_bias", False)), o_bias=False, mlp_bias=False)


class NeuronStableLmForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronStableLmModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        return fuse_qkv_and_gate_up(sd, config.num_hidden_layers)


# ---- Cohere -------------------------------------------------------------------------------------------------------------
class NeuronCohereModel(NeuronClassicModel):
    SPEC = dict(NeuronClassicModel.SPEC, norm_bias=False)

    def layer_spec(self, config, i):
        b = bool(getattr(config, "attention_bias", False))
        return dict(parallel=True, shared_norm=True, norm_bias=False, mlp="gated", act=getattr(config, "hidden_act", "silu"), qkv_bias=b,
                    o_bias=b, mlp_bias=False, rope_interleaved=True)


class NeuronCohereForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronCohereModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        return fuse_qkv_and_gate_up(sd, config.num_hidden_layers)


# ---- GPT-NeoX / Pythia ---------------------------------------------------------------------------------------------------
class NeuronGPTNeoXModel(NeuronClassicModel):
    def layer_spec(self, config, i):
        return dict(parallel=bool(getattr(config, "use_parallel_residual", True)), norm_bias=True, mlp="plain",
                    act=getattr(config, "hidden_act", "gelu"), qkv_bias=bool(getattr(config, "attention_bias", True)),
                    o_bias=bool(getattr(config, "attention_bias", True)), mlp_bias=True)


class NeuronGPTNeoXForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronGPTNeoXModel
    _STATE_DICT_MODEL_PREFIX = "gpt_neox."

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        nh, D = config.num_attention_heads, config.hidden_size // config.num_attention_heads
        for k, v in sd.items():
            if k.endswith("attention.query_key_value.weight") or k.endswith("attention.query_key_value.bias"):
                # [heads, (q,k,v), D, ...] -> [q heads; k heads; v heads]
                w = v.view(nh, 3, D, *v.shape[1:])
                v = torch.cat([w[:, j].reshape(nh * D, *v.shape[1:]) for j in range(3)], 0)
                k = k.replace("attention.query_key_value", "self_attn.qkv_proj")
            k = (k.replace("attention.dense.", "self_attn.o_proj.").replace("mlp.dense_h_to_4h.", "mlp.fc1.").replace("mlp.dense_4h_to_h.", "mlp.fc2.")
                 .replace("embed_in.", "embed_tokens.").replace("final_layer_norm.", "norm.").replace("embed_out.", "lm_head."))
            if "rotary_emb" in k or "masked_bias" in k or k.endswith("attention.bias"):
                continue
            out[k] = v
        return out


# ---- GPT-2 ---------------------------------------------------------------------------------------------------------------
class GPT2InferenceConfig(ClassicInferenceConfig):
    attribute_map = {"n_embd": "hidden_size", "n_head": "num_attention_heads", "n_layer": "num_hidden_layers", "n_positions": "max_position_embeddings",
                     "n_inner": "intermediate_size"}

    def get_required_attributes(self):
        return ["hidden_size", "num_attention_heads", "num_hidden_layers", "vocab_size"]


class NeuronGPT2Model(NeuronClassicModel):
    learned_positions = True

    def layer_spec(self, config, i):
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "activation_function", "gelu_new"), qkv_bias=True, o_bias=True,
                    mlp_bias=True)


class NeuronGPT2ForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronGPT2Model
    _STATE_DICT_MODEL_PREFIX = "transformer."

    @classmethod
    def get_config_cls(cls):
        return GPT2InferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"):
                continue
            if any(k.endswith(s) for s in ("c_attn.weight", "attn.c_proj.weight", "c_fc.weight", "mlp.c_proj.weight")):
                v = v.t().contiguous()                      # Conv1D stores [in, out]
            k = (k.replace("h.", "layers.", 1) if k.startswith("h.") else k)
            k = (k.replace(".attn.c_attn.", ".self_attn.qkv_proj.").replace(".attn.c_proj.", ".self_attn.o_proj.").replace(".mlp.c_fc.", ".mlp.fc1.")
                 .replace(".mlp.c_proj.", ".mlp.fc2.").replace(".ln_1.", ".input_layernorm.").replace(".ln_2.", ".post_attention_layernorm."))
            k = k.replace("wte.", "embed_tokens.").replace("wpe.", "embed_positions.").replace("ln_f.", "norm.")
            out[k] = v
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


# ---- OPT -----------------------------------------------------------------------------------------------------------------
class OPTInferenceConfig(ClassicInferenceConfig):
    def add_derived_config(self):
        self.intermediate_size = getattr(self, "ffn_dim", 4 * self.hidden_size)
        super().add_derived_config()


class NeuronOPTModel(NeuronClassicModel):
    learned_positions = True
    position_offset = 2          # OPT reserves the first two rows of its position table

    def layer_spec(self, config, i):
        b = bool(getattr(config, "enable_bias", True))
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "activation_function", "relu"), qkv_bias=b, o_bias=b, mlp_bias=b)


class NeuronOPTForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronOPTModel

    @classmethod
    def get_config_cls(cls):
        return OPTInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = {(k[len("decoder."):] if k.startswith("decoder.") else k): v for k, v in sd.items()}
        out = {}
        for k, v in sd.items():
            if k.startswith("layers."):
                k = (k.replace(".self_attn.out_proj.", ".self_attn.o_proj.").replace(".self_attn_layer_norm.", ".input_layernorm.")
                     .replace(".final_layer_norm.", ".post_attention_layernorm.").replace(".fc1.", ".mlp.fc1.").replace(".fc2.", ".mlp.fc2."))
            elif k.startswith("final_layer_norm."):
                k = k.replace("final_layer_norm.", "norm.")
            out[k] = v
        out = fuse_qkv_and_gate_up(out, config.num_hidden_layers, fuse_mlp=False)
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


# ---- GPT-J ---------------------------------------------------------------------------------------------------------------
class GPTJInferenceConfig(GPT2InferenceConfig):
    pass


class NeuronGPTJModel(NeuronClassicModel):
    lm_head_bias = True

    def make_rotary(self, config, device):
        rot = int(getattr(config, "rotary_dim", None) or config.head_dim)
        return RotaryEmbedding(rot, max(config.max_position_embeddings, config.neuron_config.seq_len), rope_theta_of(config), None, device=device)

    def layer_spec(self, config, i):
        return dict(parallel=True, shared_norm=True, norm_bias=True, mlp="plain", act=getattr(config, "activation_function", "gelu_new"),
                    qkv_bias=False, o_bias=False, mlp_bias=True, rope_interleaved=True)


class NeuronGPTJForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronGPTJModel
    _STATE_DICT_MODEL_PREFIX = "transformer."

    @classmethod
    def get_config_cls(cls):
        return GPTJInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"):
                continue
            k = k.replace("h.", "layers.", 1) if k.startswith("h.") else k
            k = (k.replace(".attn.q_proj.", ".self_attn.q_proj.").replace(".attn.k_proj.", ".self_attn.k_proj.").replace(".attn.v_proj.", ".self_attn.v_proj.")
                 .replace(".attn.out_proj.", ".self_attn.o_proj.").replace(".mlp.fc_in.", ".mlp.fc1.").replace(".mlp.fc_out.", ".mlp.fc2.")
                 .replace(".ln_1.", ".input_layernorm."))
            k = k.replace("wte.", "embed_tokens.").replace("ln_f.", "norm.")
            out[k] = v
        return fuse_qkv_and_gate_up(out, config.num_hidden_layers, fuse_mlp=False)


# ---- Phi-1 / Phi-1.5 / Phi-2 -------------------------------------------------------------------------------------------------
class NeuronPhiModel(NeuronClassicModel):
    lm_head_bias = True

    def layer_spec(self, config, i):
        return dict(parallel=True, shared_norm=True, norm_bias=True, mlp="plain", act=getattr(config, "hidden_act", "gelu_new"), qkv_bias=True,
                    o_bias=True, mlp_bias=True)


class NeuronPhiForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronPhiModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = {k.replace(".self_attn.dense.", ".self_attn.o_proj.").replace("final_layernorm.", "norm."): v for k, v in sd.items()}
        return fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)


# ---- Falcon (7B: multi-query, one LayerNorm; 40B / 180B / Falcon2: grouped-query, per-branch LayerNorms), parallel attention + MLP ---------------------------------------------------
class FalconInferenceConfig(ClassicInferenceConfig):
    def add_derived_config(self):
        if getattr(self, "new_decoder_architecture", False):
            self.num_key_value_heads = getattr(self, "num_kv_heads", self.num_attention_heads)
        else:
            self.num_key_value_heads = 1 if getattr(self, "multi_query", True) else self.num_attention_heads
        self.intermediate_size = getattr(self, "ffn_hidden_size", None) or 4 * self.hidden_size
        super().add_derived_config()

    def validate_config(self):
        super().validate_config()
        if getattr(self, "alibi", False) or not (getattr(self, "parallel_attn", True) or getattr(self, "new_decoder_architecture", False)):
            raise NotImplementedError("Falcon: the rotary, parallel-attention layouts (7B-style and the 40B / 180B / Falcon2 decoder) are implemented")


class NeuronFalconModel(NeuronClassicModel):
    def layer_spec(self, config, i):
        b = bool(getattr(config, "bias", False))
        # 40B / 180B decoder: grouped-query attention and separate LayerNorms in front of the attention and the MLP branches
        two_ln = bool(getattr(config, "new_decoder_architecture", False)) and (getattr(config, "num_ln_in_parallel_attn", None) or 2) == 2
        return dict(parallel=True, shared_norm=not two_ln, norm_bias=True, mlp="plain", act="gelu", qkv_bias=b, o_bias=b, mlp_bias=b)


class NeuronFalconForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronFalconModel
    _STATE_DICT_MODEL_PREFIX = "transformer."

    @classmethod
    def get_config_cls(cls):
        return FalconInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            k = k.replace("h.", "layers.", 1) if k.startswith("h.") else k
            k = (k.replace(".self_attention.query_key_value.", ".self_attn.qkv_proj.").replace(".self_attention.dense.", ".self_attn.o_proj.")
                 .replace(".mlp.dense_h_to_4h.", ".mlp.fc1.").replace(".mlp.dense_4h_to_h.", ".mlp.fc2."))
            k = k.replace("word_embeddings.", "embed_tokens.").replace("ln_f.", "norm.")
            k = k.replace(".ln_attn.", ".input_layernorm.").replace(".ln_mlp.", ".post_attention_layernorm.")
            if ".self_attn.qkv_proj." in k and getattr(config, "new_decoder_architecture", False):
                # rows are [kv group, (its q heads..., k, v), D]  ->  [all q heads; all k; all v]
                nkv, D = config.num_key_value_heads, config.hidden_size // config.num_attention_heads
                w = v.view(nkv, config.num_attention_heads // nkv + 2, D, *v.shape[1:])
                v = torch.cat([w[:, :-2].reshape(-1, *v.shape[1:]), w[:, -2].reshape(-1, *v.shape[1:]), w[:, -1].reshape(-1, *v.shape[1:])], 0)
            out[k] = v        # (7B-style multi-query fused QKV is already [q heads; k; v])
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


# ---- GPT-BigCode (SantaCoder / StarCoder-1): GPT-2 block with multi-query attention and nn.Linear weights --------------------------
class GPTBigCodeInferenceConfig(GPT2InferenceConfig):
    def add_derived_config(self):
        self.num_key_value_heads = 1 if getattr(self, "multi_query", True) else self.num_attention_heads
        super().add_derived_config()


class NeuronGPTBigCodeModel(NeuronClassicModel):
    learned_positions = True

    def layer_spec(self, config, i):
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "activation_function", "gelu_pytorch_tanh"), qkv_bias=True,
                    o_bias=True, mlp_bias=True)


class NeuronGPTBigCodeForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronGPTBigCodeModel
    _STATE_DICT_MODEL_PREFIX = "transformer."

    @classmethod
    def get_config_cls(cls):
        return GPTBigCodeInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"):
                continue
            k = k.replace("h.", "layers.", 1) if k.startswith("h.") else k
            k = (k.replace(".attn.c_attn.", ".self_attn.qkv_proj.").replace(".attn.c_proj.", ".self_attn.o_proj.").replace(".mlp.c_fc.", ".mlp.fc1.")
                 .replace(".mlp.c_proj.", ".mlp.fc2.").replace(".ln_1.", ".input_layernorm.").replace(".ln_2.", ".post_attention_layernorm."))
            out[k.replace("wte.", "embed_tokens.").replace("wpe.", "embed_positions.").replace("ln_f.", "norm.")] = v
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


# ---- GPT-Neo: unscaled attention logits, alternating global / local (windowed) layers ----------------------------------------------
class GPTNeoInferenceConfig(ClassicInferenceConfig):
    attribute_map = {"num_layers": "num_hidden_layers", "num_heads": "num_attention_heads"}

    def add_derived_config(self):
        if getattr(self, "intermediate_size", None) is None:
            self.intermediate_size = 4 * self.hidden_size
        super().add_derived_config()


class NeuronGPTNeoModel(NeuronClassicModel):
    learned_positions = True

    def layer_spec(self, config, i):
        kinds = getattr(config, "attention_layers", None)
        if not kinds:
            kinds = []
            for pattern, n in getattr(config, "attention_types", [[["global"], config.num_hidden_layers]]):
                kinds += list(pattern) * n
        local = kinds[i] == "local"
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "activation_function", "gelu_new"), qkv_bias=False, o_bias=True,
                    mlp_bias=True, softmax_scale=1.0, sliding_window=getattr(config, "window_size", 256) if local else None)


class NeuronGPTNeoForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronGPTNeoModel
    _STATE_DICT_MODEL_PREFIX = "transformer."

    @classmethod
    def get_config_cls(cls):
        return GPTNeoInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k.endswith(".attention.bias") or k.endswith(".attention.masked_bias"):
                continue
            k = k.replace("h.", "layers.", 1) if k.startswith("h.") else k
            k = (k.replace(".attn.attention.out_proj.", ".self_attn.o_proj.").replace(".attn.attention.", ".self_attn.").replace(".mlp.c_fc.", ".mlp.fc1.")
                 .replace(".mlp.c_proj.", ".mlp.fc2.").replace(".ln_1.", ".input_layernorm.").replace(".ln_2.", ".post_attention_layernorm."))
            out[k.replace("wte.", "embed_tokens.").replace("wpe.", "embed_positions.").replace("ln_f.", "norm.")] = v
        out = fuse_qkv_and_gate_up(out, config.num_hidden_layers, fuse_mlp=False)
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


# ---- BioGPT: OPT-style block, sqrt(H)-scaled embeddings -------------------------------------------------------------------------------
class NeuronBioGptModel(NeuronClassicModel):
    learned_positions = True
    position_offset = 2

    def init_model(self, config):
        super().init_model(config)
        if getattr(config, "scale_embedding", True):
            self.embed_scale = float(config.hidden_size ** 0.5)

    def embed(self, input_ids, inputs_embeds=None, vision_embeddings=None, vision_mask=None):
        h = self.embed_tokens(input_ids) * getattr(self, "embed_scale", 1.0)
        pos = self._pos if self._pos is not None else torch.arange(input_ids.shape[1], device=input_ids.device).unsqueeze(0)
        return h + self.embed_positions((pos.long() + self.position_offset).clamp(0, self.embed_positions.num_embeddings - 1))

    def layer_spec(self, config, i):
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "hidden_act", "gelu"), qkv_bias=True, o_bias=True, mlp_bias=True)


class NeuronBioGptForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronBioGptModel
    _STATE_DICT_MODEL_PREFIX = "biogpt."

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k.startswith("layers."):
                k = (k.replace(".self_attn.out_proj.", ".self_attn.o_proj.").replace(".self_attn_layer_norm.", ".input_layernorm.")
                     .replace(".final_layer_norm.", ".post_attention_layernorm.").replace(".fc1.", ".mlp.fc1.").replace(".fc2.", ".mlp.fc2."))
            elif k.startswith("layer_norm."):
                k = k.replace("layer_norm.", "norm.")
            elif k.startswith("output_projection."):
                k = k.replace("output_projection.", "lm_head.")
            out[k] = v
        out = fuse_qkv_and_gate_up(out, config.num_hidden_layers, fuse_mlp=False)
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


CLASSIC_MODEL_TYPES = {"gpt_bigcode": NeuronGPTBigCodeForCausalLM, "gpt_neo": NeuronGPTNeoForCausalLM, "biogpt": NeuronBioGptForCausalLM,
                       "opt": NeuronOPTForCausalLM, "gptj": NeuronGPTJForCausalLM, "phi": NeuronPhiForCausalLM, "falcon": NeuronFalconForCausalLM,"starcoder2": NeuronStarcoder2ForCausalLM, "stablelm": NeuronStableLmForCausalLM, "cohere": NeuronCohereForCausalLM,
                       "gpt_neox": NeuronGPTNeoXForCausalLM, "gpt2": NeuronGPT2ForCausalLM}
