"""Llama-4 text decoder (reference models/llama4/modeling_llama4_text.py:1-770).

Per layer: RoPE layers use *interleaved* (complex) rotary, weight-less L2 q/k norm after RoPE and chunked causal
attention (``attention_chunk_size``); every ``nope`` layer (``no_rope_layers[i] == 0``) has no rotary, full causal
attention and log-position temperature tuning of q.  Feed-forward: dense SwiGLU (``intermediate_size_mlp``) or, on
``moe_layers``, a sigmoid top-k router whose affinity scales the expert *input* plus an always-on shared expert."""
from __future__ import annotations

import torch

from ...config import MoENeuronConfig
from ...modules.mlp import GatedMLP
from ...modules.moe import initialize_moe_module
from ...modules.norm import RMSNorm
from ..llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel
from ..model_base import DecoderLayer
from ..state_dict_utils import fuse_qkv_and_gate_up


class Llama4InferenceConfig(LlamaInferenceConfig):
    def add_derived_config(self):
        tc = getattr(self, "text_config", None)
        if tc is not None and not hasattr(self, "hidden_size"):
            for k, v in vars(tc).items():
                if k != "neuron_config" and not hasattr(self, k):
                    setattr(self, k, v)
        super().add_derived_config()
        n = self.num_hidden_layers
        if not getattr(self, "no_rope_layers", None):
            step = getattr(self, "no_rope_layer_interval", 4)
            self.no_rope_layers = [int((i + 1) % step != 0) for i in range(n)]
        if not getattr(self, "moe_layers", None):
            step = getattr(self, "interleave_moe_layer_step", 1)
            self.moe_layers = list(range(step - 1, n, step))

    @classmethod
    def get_neuron_config_cls(cls):
        return MoENeuronConfig


class NeuronLlama4Attention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None):
        use_rope = bool(config.no_rope_layers[layer_idx])
        super().__init__(config, layer_idx, rotary_emb, device=device, use_rope=use_rope, rope_interleaved=True,
                         qk_norm="l2_post_rope" if (getattr(config, "use_qk_norm", False) and use_rope) else None,
                         qk_norm_eps=config.rms_norm_eps,
                         attention_chunk_size=getattr(config, "attention_chunk_size", None) if use_rope else None)
        self.temperature_tuning = bool(getattr(config, "attn_temperature_tuning", False)) and not use_rope
        self.attn_scale = float(getattr(config, "attn_scale", 0.1))
        self.floor_scale = float(getattr(config, "floor_scale", 8192))

    def _split_norm_rope(self, qkv, B, T, cos, sin, meta=None):
        q, k, v = super()._split_norm_rope(qkv, B, T, cos, sin, meta)
        if self.temperature_tuning and meta is not None:
            pos = meta.position_ids.float()
            s = torch.log1p(torch.floor((pos + 1.0) / self.floor_scale)) * self.attn_scale + 1.0
            q = (q * s.view(B, T, 1, 1)).to(q.dtype)
        return q, k, v


class NeuronLlama4TextModel(NeuronLlamaModel):
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        nc = config.neuron_config
        dt = nc.torch_dtype
        attn = NeuronLlama4Attention(config, i, rotary, device=device)
        if i in config.moe_layers:
            mlp = initialize_moe_module(config, device=device, router_act="sigmoid", apply_act_fn_over_topk=True, shared=True,
                                        early_affinity_modulation=True, intermediate_size=config.intermediate_size)
        else:
            mlp = GatedMLP(config.hidden_size, getattr(config, "intermediate_size_mlp", config.intermediate_size),
                           config.hidden_act, dt, device=device)
        return DecoderLayer(attn, mlp, RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device), i)


class NeuronLlama4TextForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronLlama4TextModel

    @classmethod
    def get_config_cls(cls):
        return Llama4InferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(state_dict, config):
        sd = {k.replace("language_model.model.", "").replace("language_model.", ""): v for k, v in state_dict.items()
              if "vision_model" not in k and "multi_modal_projector" not in k}
        sd = {k.replace(".feed_forward.", ".mlp."): v for k, v in sd.items()}
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers)
        out = {}
        for k, v in sd.items():
            if k.endswith(".mlp.router.weight"):
                out[k.replace(".mlp.router.weight", ".mlp.router.linear_router.weight")] = v.float()
            elif k.endswith(".mlp.experts.gate_up_proj"):       # [E, H, 2I] (gate half | up half)
                out[k.replace(".mlp.experts.gate_up_proj", ".mlp.expert_mlps.gate_up_proj")] = v.transpose(1, 2).contiguous()
            elif k.endswith(".mlp.experts.down_proj"):           # [E, I, H]
                out[k.replace(".mlp.experts.down_proj", ".mlp.expert_mlps.down_proj")] = v.transpose(1, 2).contiguous()
            elif ".mlp.shared_expert." in k:
                k2 = k.replace(".mlp.shared_expert.", ".mlp.shared_experts.")
                out[k2] = v
            else:
                out[k] = v
        # the generic pass fused shared_expert gate/up only under ".mlp." prefix; fuse the shared expert explicitly
        for i in config.moe_layers:
            g, u = f"layers.{i}.mlp.shared_experts.gate_proj.weight", f"layers.{i}.mlp.shared_experts.up_proj.weight"
            if g in out:
                out[f"layers.{i}.mlp.shared_experts.gate_up_proj.weight"] = torch.cat([out.pop(g), out.pop(u)], 0)
        return out
