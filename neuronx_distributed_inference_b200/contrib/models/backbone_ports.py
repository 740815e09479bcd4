"""Ports that are an existing decoder behind a different checkpoint / config layout.

* **MiniCPM4** — Llama block with muP scaling: ``scale_emb`` on the embeddings, ``scale_depth / sqrt(L)`` on both residual branches,
  hidden states divided by ``hidden_size / dim_model_base`` before the head  ==  Granite's four multipliers.
* **InternLM3** — Llama with separately switchable q/k/v bias (``qkv_bias``) and o_proj / MLP bias (``bias``).
* **Orion** — Llama layout with LayerNorm (+bias) instead of RMSNorm.
* **Janus / Ovis2.5 / Qwen2.5-Omni thinker (text backbones)** — Llama / Qwen3 / Qwen2.5 decoders nested in a multimodal checkpoint (``language_model.*`` / ``llm.*``
  weights, ``text_config`` / ``llm_config`` hyper-parameters); like the reference ports, only the language path is served.
reference ports: contrib/models/{MiniCPM4-8B, internlm3-8b-instruct, orion-14b-chat, Janus-1.3B, Ovis2.5-9B, Qwen2.5-Omni-7B}/src."""
from __future__ import annotations

import math

from ...models.llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel
from ...models.qwen2.modeling_qwen2 import NeuronQwen2ForCausalLM
from ...models.qwen3.modeling_qwen3 import NeuronQwen3ForCausalLM
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from ...modules.mlp import GatedMLP
from .classic_family import ClassicInferenceConfig, NeuronClassicModel, _ClassicCausalLM
from .llama_family import NeuronGraniteForCausalLM


# ---------------------------------------------------------------------------------------------------------------------- MiniCPM4
class MiniCPMInferenceConfig(LlamaInferenceConfig):
    def add_derived_config(self):
        super().add_derived_config()
        self.embedding_multiplier = float(getattr(self, "scale_emb", 1.0))
        self.residual_multiplier = float(getattr(self, "scale_depth", 1.0)) / math.sqrt(self.num_hidden_layers)
        self.logits_scaling = self.hidden_size / float(getattr(self, "dim_model_base", self.hidden_size))
        self.attention_multiplier = 1.0 / math.sqrt(self.head_dim)


class NeuronMiniCPMForCausalLM(NeuronGraniteForCausalLM):
    @classmethod
    def get_config_cls(cls):
        return MiniCPMInferenceConfig


# ---------------------------------------------------------------------------------------------------------------------- InternLM3
class _InternLM3Attention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, qkv_bias=bool(getattr(config, "qkv_bias", False)),
                         o_bias=bool(getattr(config, "bias", False)), **over)


class _InternLM3MLP(GatedMLP):
    def __init__(self, config, device=None):
        nc = config.neuron_config
        super().__init__(config.hidden_size, config.intermediate_size, config.hidden_act, nc.torch_dtype, bias=bool(getattr(config, "bias", False)),
                         device=device, sequence_parallel_enabled=nc.sequence_parallel_enabled, reduce_dtype=nc.rpl_reduce_dtype)


class NeuronInternLM3Model(NeuronLlamaModel):
    attention_cls = _InternLM3Attention
    mlp_cls = _InternLM3MLP


class NeuronInternLM3ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronInternLM3Model


# ---------------------------------------------------------------------------------------------------------------------- Orion
class NeuronOrionModel(NeuronClassicModel):
    def layer_spec(self, config, i):
        b = bool(getattr(config, "attention_bias", False))
        return dict(parallel=False, norm_bias=True, mlp="gated", act=getattr(config, "hidden_act", "silu"), qkv_bias=b, o_bias=b, mlp_bias=False)


class NeuronOrionForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronOrionModel

    @classmethod
    def get_config_cls(cls):
        return ClassicInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        return fuse_qkv_and_gate_up(sd, config.num_hidden_layers)


# ---------------------------------------------------------------------------------------------------------------------- text backbones
_TEXT_PREFIXES = ("thinker.model.", "thinker.", "model.language_model.", "language_model.model.", "language_model.", "llm.model.", "llm.")
_TEXT_CONFIG_KEYS = ("text_config", "llm_config", "language_config")


def _hoist_text_config(cfg):
    """Copy the nested text hyper-parameters to the top level (only names the top level does not define)."""
    thinker = getattr(cfg, "thinker_config", None)           # Qwen2.5-Omni: thinker_config.text_config
    if thinker is not None:
        sub = thinker.get("text_config") if isinstance(thinker, dict) else getattr(thinker, "text_config", None)
        if sub is not None:
            object.__setattr__(cfg, "text_config", sub)
    for key in _TEXT_CONFIG_KEYS:
        sub = getattr(cfg, key, None)
        if sub is None:
            continue
        items = sub.items() if isinstance(sub, dict) else ((k, v) for k, v in vars(sub).items() if k != "neuron_config")
        for k, v in items:
            if not hasattr(cfg, k) or getattr(cfg, k) is None or k in ("vocab_size", "hidden_size", "tie_word_embeddings"):
                setattr(cfg, k, v)
        return


def _text_backbone(base_cls, name):
    base_cfg = base_cls.get_config_cls()

    class Cfg(base_cfg):
        def add_derived_config(self):
            _hoist_text_config(self)
            super().add_derived_config()

        def validate_config(self):
            _hoist_text_config(self)
            super().validate_config()

    class Port(base_cls):
        @classmethod
        def get_config_cls(cls):
            return Cfg

        @classmethod
        def _strip(cls, k):
            for p in _TEXT_PREFIXES:
                if k.startswith(p):
                    return k[len(p):]
            return super()._strip(k)

        @staticmethod
        def update_state_dict_for_tied_weights(sd):
            if "lm_head.weight" not in sd:       # composite configs often say "tied" while the checkpoint ships its own head
                sd["lm_head.weight"] = sd["embed_tokens.weight"].clone()

        @classmethod
        def get_state_dict(cls, path, config):
            sd = super().get_state_dict(path, config)
            keep = ("embed_tokens.", "layers.", "norm.", "lm_head.")
            return {k: v for k, v in sd.items() if k.startswith(keep)}      # vision / generation heads are not served

    Cfg.__name__, Port.__name__ = f"{name}InferenceConfig", f"Neuron{name}ForCausalLM"
    return Port


NeuronJanusForCausalLM = _text_backbone(NeuronLlamaForCausalLM, "Janus")
NeuronOvis2_5ForCausalLM = _text_backbone(NeuronQwen3ForCausalLM, "Ovis2_5")
NeuronQwen2_5OmniForCausalLM = _text_backbone(NeuronQwen2ForCausalLM, "Qwen2_5Omni")      # text-only prompts: M-RoPE == RoPE

PORT_MODEL_TYPES = {"minicpm": NeuronMiniCPMForCausalLM, "internlm3": NeuronInternLM3ForCausalLM, "orion": NeuronOrionForCausalLM,
                    "janus": NeuronJanusForCausalLM, "ovis2_5": NeuronOvis2_5ForCausalLM, "qwen2_5_omni": NeuronQwen2_5OmniForCausalLM}
