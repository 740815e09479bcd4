"""Mllama (Llama-3.2 Vision): tiled ViT (local + gated global transformer, gated tile / position embeddings) whose output is
consumed by *cross-attention* decoder layers interleaved with the Llama self-attention layers.

reference: models/mllama/* (≈3380 LoC: modeling_mllama.py, modeling_mllama_vision.py, its own model wrapper and a vision-token
KV cache).  Here the cross-attention K/V of every cross layer are projected once at prefill and kept in a per-layer buffer
indexed by cache line (``seq_ids``), exactly like the self-attention cache; decode steps reuse them, with the visibility row of
the last prompt token (HF generation semantics)."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from ...modules.gqa import GroupQueryAttention_O, GroupQueryAttention_QKV
from ...modules.norm import RMSNorm
from ...modules.vision import PatchEmbed, VisionAttention, VisionMLP
from ...parallel.layers import ColumnParallelLinear, ParallelEmbedding
from ..image_to_text_model_base import ImageToTextInferenceConfig, NeuronBaseForImageToText
from ..llama.modeling_llama import NeuronLlamaMLP, NeuronLlamaModel
from ..state_dict_utils import fuse_qkv_and_gate_up


class MllamaInferenceConfig(ImageToTextInferenceConfig):
    def get_required_attributes(self):
        return ["text_config", "vision_config"]


# ---------------------------------------------------------------------------------------------------------------------
# vision tower
class MllamaVisionLayer(nn.Module):
    def __init__(self, vc, gated: bool, dtype, device):
        super().__init__()
        self.input_layernorm = nn.LayerNorm(vc.hidden_size, eps=getattr(vc, "norm_eps", 1e-5), dtype=dtype, device=device)
        self.post_attention_layernorm = nn.LayerNorm(vc.hidden_size, eps=getattr(vc, "norm_eps", 1e-5), dtype=dtype, device=device)
        self.self_attn = VisionAttention(vc.hidden_size, vc.attention_heads, False, dtype, device)
        self.mlp = VisionMLP(vc.hidden_size, vc.intermediate_size, getattr(vc, "hidden_act", "gelu"), True, False, dtype, device)
        self.gated = gated
        if gated:
            self.gate_attn = nn.Parameter(torch.zeros(1, dtype=dtype, device=device), requires_grad=False)
            self.gate_ffn = nn.Parameter(torch.zeros(1, dtype=dtype, device=device), requires_grad=False)

    def forward(self, x, mask):
        a = self.self_attn(self.input_layernorm(x), mask=mask)
        x = x + (self.gate_attn.tanh() * a if self.gated else a)
        m = self.mlp(self.post_attention_layernorm(x))
        return x + (self.gate_ffn.tanh() * m if self.gated else m)


class NeuronMllamaVisionModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        vc = config.vision_config
        dt = vc.neuron_config.torch_dtype
        self.vc = vc
        H, T = vc.hidden_size, vc.max_num_tiles
        self.num_patches = (vc.image_size // vc.patch_size) ** 2 + 1
        n_ar = len(vc.supported_aspect_ratios) + 1
        self.patch_embedding = PatchEmbed(vc.num_channels * vc.patch_size ** 2, H, False, dt, device)
        P = lambda *s: nn.Parameter(torch.zeros(*s, dtype=dt, device=device), requires_grad=False)  # noqa: E731
        self.class_embedding = P(H)
        self.gated_positional_embedding = nn.ParameterDict({"gate": P(1), "embedding": P(self.num_patches, H)})
        self.gated_positional_tile_embedding = nn.Embedding(n_ar, T * self.num_patches * H, dtype=dt, device=device)
        self.pre_tile_gate, self.post_tile_gate = P(1), P(1)
        self.pre_tile_embedding = nn.Embedding(n_ar, T * H, dtype=dt, device=device)
        self.post_tile_embedding = nn.Embedding(n_ar, T * H, dtype=dt, device=device)
        self.layernorm_pre = nn.LayerNorm(H, dtype=dt, device=device)
        self.layernorm_post = nn.LayerNorm(H, dtype=dt, device=device)
        self.layers = nn.ModuleList([MllamaVisionLayer(vc, False, dt, device) for _ in range(vc.num_hidden_layers)])
        self.global_layers = nn.ModuleList([MllamaVisionLayer(vc, True, dt, device) for _ in range(vc.num_global_layers)])
        self.inter = list(vc.intermediate_layers_indices)
        # Meta's implementation and transformers 4.x export the INPUT of layer i for index i (what released checkpoints were
        # tuned with); transformers 5.x returns the OUTPUT of layer i.  Default: the original semantics.
        self.intermediate_is_layer_output = bool(getattr(vc, "intermediate_is_layer_output", False))
        self.projector = nn.Linear(vc.vision_output_dim, config.get_text_config().hidden_size, dtype=dt, device=device)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, pixel_values, aspect_ratio_ids, aspect_ratio_mask):
        B, M, T, C, Hh, Ww = pixel_values.shape
        vc, P, H = self.vc, self.vc.patch_size, self.vc.hidden_size
        x = pixel_values.reshape(B * M * T, C, Hh // P, P, Ww // P, P).permute(0, 2, 4, 1, 3, 5).reshape(B * M * T, -1, C * P * P)
        x = self.patch_embedding(x)                                            # [BMT, np, H]
        ar = aspect_ratio_ids.reshape(B * M)
        n0 = x.shape[1]
        x = x.view(B * M, T, n0, H) + (self.pre_tile_embedding(ar).view(B * M, T, 1, H) * self.pre_tile_gate.tanh())
        x = torch.cat([self.class_embedding.view(1, 1, 1, H).expand(B * M, T, 1, H), x], 2)
        N = n0 + 1
        g = self.gated_positional_embedding["gate"].tanh()
        x = x + (1 - g) * self.gated_positional_embedding["embedding"].view(1, 1, N, H)
        x = x + g * self.gated_positional_tile_embedding(ar).view(B * M, T, N, H)
        x = self.layernorm_pre(x)
        pad = (8 - N % 8) % 8
        if pad:
            x = torch.cat([x, x.new_zeros(B * M, T, pad, H)], 2)
        L = N + pad
        # HF masks only (invalid, invalid) pairs: padded tiles / padding patches still exchange with real ones
        valid = aspect_ratio_mask.reshape(B * M, T, 1).to(x.dtype).repeat(1, 1, L)
        valid[:, :, N:] = 0
        inv = (1 - valid).reshape(B * M, T * L)
        mask = ~(inv.unsqueeze(-1) * inv.unsqueeze(-2)).bool().unsqueeze(1)     # True = may attend
        x = x.view(B * M, T * L, H)
        hs = []
        for layer in self.layers:
            hs.append(x)
            x = layer(x, mask)
        hs.append(x)
        x = self.layernorm_post(x)
        x = x.view(B * M, T, L, H) + self.post_tile_embedding(ar).view(B * M, T, 1, H) * self.post_tile_gate.tanh()
        x = x.view(B * M, T * L, H)
        for layer in self.global_layers:
            x = layer(x, mask)
        x = x.view(B * M, T, L, H)[:, :, :N]
        off = 1 if self.intermediate_is_layer_output else 0
        inter = torch.stack([hs[i + off] for i in self.inter], -1).view(B * M, T, L, -1)[:, :, :N]
        feat = torch.cat([x, inter], -1).view(B, M * T * N, -1)
        return self.projector(feat)                                             # [B, M*T*N, H_text]


# ---------------------------------------------------------------------------------------------------------------------
# text decoder
class MllamaCrossAttentionLayer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, layer_idx: int, device=None):
        super().__init__()
        nc = config.neuron_config
        dt = nc.torch_dtype
        D = config.head_dim
        self.layer_idx, self.head_dim = layer_idx, D
        self.qkv_proj = GroupQueryAttention_QKV(config.hidden_size, D, config.num_attention_heads, config.num_key_value_heads,
                                                None, dt, False, None, device)
        self.o_proj = GroupQueryAttention_O(config.hidden_size, D, config.num_attention_heads, config.num_key_value_heads,
                                            None, dt, False, None, device)
        self.n_q, self.n_kv = self.qkv_proj.n_q, self.qkv_proj.n_kv
        self.q_norm = RMSNorm(D, config.rms_norm_eps, dt, device=device)
        self.k_norm = RMSNorm(D, config.rms_norm_eps, dt, device=device)
        self.input_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.post_attention_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.mlp = NeuronLlamaMLP(config, device=device)
        self.cross_attn_attn_gate = nn.Parameter(torch.zeros(1, dtype=dt, device=device), requires_grad=False)
        self.cross_attn_mlp_gate = nn.Parameter(torch.zeros(1, dtype=dt, device=device), requires_grad=False)

    def forward(self, h, meta, kv_mgr, lora=None):
        """``kv_mgr``: a MultimodalKVCacheManager — holds this layer's vision K/V per cache line."""
        states = meta.extras.get("cross_attention_states")
        if states is None and not kv_mgr.has_vision(self.layer_idx):
            return h                                                             # text-only request: layer is skipped
        B, T, _ = h.shape
        D, nq, nkv = self.head_dim, self.n_q, self.n_kv
        lines = kv_mgr.lines_for(meta.seq_ids)
        w = self.qkv_proj.weight
        x = self.input_layernorm(h)
        q = self.q_norm(ops.linear(x, w[: nq * D]).view(B, T, nq, D))
        if states is not None:
            Nv = states.shape[1]
            kv = ops.linear(states.to(h.dtype), w[nq * D:]).view(B, Nv, 2 * nkv, D)
            k, v = self.k_norm(kv[:, :, :nkv]).transpose(1, 2), kv[:, :, nkv:].transpose(1, 2)
            cm = meta.extras.get("cross_attention_mask")                        # [B, T0, Nv] bool
            if cm is None:
                cm = torch.ones(B, T, Nv, dtype=torch.bool, device=h.device)
            if cm.shape[1] < T:
                cm = torch.cat([cm, cm.new_zeros(B, T - cm.shape[1], Nv)], 1)
            last = (meta.key_valid.long().sum(-1).clamp_min(1) - 1) if meta.key_valid is not None else \
                torch.full((B,), T - 1, device=h.device)
            kv_mgr.update_vision(self.layer_idx, lines, k, v, cm[torch.arange(B, device=h.device), last])
        else:
            k, v, rm = kv_mgr.get_vision(self.layer_idx, lines)
            cm = rm.unsqueeze(1).expand(B, T, -1)
        row_on = cm.any(-1, keepdim=True)                                        # rows that see at least one vision token
        mask = (cm | ~row_on).unsqueeze(1)                                       # fully masked rows attend uniformly (HF)
        o = ops.ref.attention_with_mask(q.transpose(1, 2), k, v, mask, 1.0 / math.sqrt(D))
        a = self.o_proj(o.transpose(1, 2).reshape(B, T, nq * D))
        h = h + self.cross_attn_attn_gate.tanh() * a
        m = self.mlp(self.post_attention_layernorm(h)) * row_on.to(h.dtype)
        return h + self.cross_attn_mlp_gate.tanh() * m


class NeuronMllamaTextModel(NeuronLlamaModel):
    meta_extra_keys = ("cross_attention_states", "cross_attention_mask")
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        if i in getattr(config, "cross_attention_layers", []):
            return MllamaCrossAttentionLayer(config, i, device)
        return super().make_layer(config, i, rotary, device)

    def init_model(self, config):
        super().init_model(config)
        nc = config.neuron_config
        self.embed_tokens = ParallelEmbedding(config.vocab_size + 8, config.hidden_size, getattr(config, "pad_token_id", None),
                                              dtype=nc.torch_dtype, device=self.device_, shard_across_embedding=not nc.vocab_parallel,
                                              pad=True, tensor_model_parallel_group=self.tp_group)

    def kv_heads_per_rank(self):
        return next(l.self_attn.n_kv for l in self.layers if hasattr(l, "self_attn"))

    def kv_head_dim(self):
        return next(l.self_attn.head_dim for l in self.layers if hasattr(l, "self_attn"))

    def init_inference_optimization(self, config):
        super().init_inference_optimization(config)
        nc = self.neuron_config
        if nc.is_block_kv_layout or nc.attention_dp_degree > 1:
            raise NotImplementedError("Mllama: contiguous KV layout only (vision K/V are stored per cache line)")
        from ...modules.kvcache.multimodal_kv_cache_manager import MultimodalKVCacheManager
        base = self.kv_mgr
        self.kv_mgr = MultimodalKVCacheManager(base.num_layers, base.num_kv_heads, base.head_dim, base.max_len, base.num_lines,
                                               nc.attention_dtype or nc.torch_dtype, self.device_,
                                               quant_config=nc.kv_quant_config if nc.kv_cache_quant else None,
                                               cross_attention_layers=[i for i, l in enumerate(self.layers)
                                                                       if isinstance(l, MllamaCrossAttentionLayer)])

    def reset(self):
        super().reset()
        self.kv_mgr.reset_vision()


class NeuronMllamaForCausalLM(NeuronBaseForImageToText):
    _model_cls = NeuronMllamaTextModel
    _vision_cls = NeuronMllamaVisionModel
    text_prefix = "language_model."
    vision_prefix = "vision_model."
    vision_kwargs = ("aspect_ratio_ids", "aspect_ratio_mask")

    @classmethod
    def get_config_cls(cls):
        return MllamaInferenceConfig

    @staticmethod
    def load_hf_model(model_path):
        from transformers import AutoModelForImageTextToText
        return AutoModelForImageTextToText.from_pretrained(model_path)

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = dict(sd)
        for i in getattr(config, "cross_attention_layers", []):
            a = f"layers.{i}.cross_attn."
            sd[f"layers.{i}.qkv_proj.weight"] = torch.cat([sd.pop(a + f"{p}_proj.weight") for p in "qkv"], 0)
            sd[f"layers.{i}.o_proj.weight"] = sd.pop(a + "o_proj.weight")
            sd[f"layers.{i}.q_norm.weight"] = sd.pop(a + "q_norm.weight")
            sd[f"layers.{i}.k_norm.weight"] = sd.pop(a + "k_norm.weight")
        return fuse_qkv_and_gate_up(sd, config.num_hidden_layers)

    @classmethod
    def get_state_dict(cls, path, config):
        from ...modules.checkpoint import load_state_dict
        sd = {cls._strip(k): v for k, v in load_state_dict(path).items()}
        text = {k[len(cls.text_prefix):] if k.startswith(cls.text_prefix) else k: v for k, v in sd.items()
                if not k.startswith(cls.vision_prefix) and not k.startswith("multi_modal_projector.")}
        text = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in text.items()}
        text = cls.convert_hf_to_neuron_state_dict(text, config.get_text_config())
        out = {cls.text_prefix + k: v for k, v in text.items()}
        vc = config.vision_config
        vis = {}
        for k, v in sd.items():
            if k.startswith("multi_modal_projector."):
                vis["projector." + k.split(".", 1)[1]] = v
            elif k.startswith(cls.vision_prefix):
                k = k[len(cls.vision_prefix):]
                k = (k.replace("global_transformer.layers.", "global_layers.").replace("transformer.layers.", "layers.")
                     .replace("gated_positional_embedding.tile_embedding.", "gated_positional_tile_embedding.")
                     .replace("pre_tile_positional_embedding.gate", "pre_tile_gate")
                     .replace("post_tile_positional_embedding.gate", "post_tile_gate")
                     .replace("pre_tile_positional_embedding.embedding.", "pre_tile_embedding.")
                     .replace("post_tile_positional_embedding.embedding.", "post_tile_embedding."))
                if k == "patch_embedding.weight":
                    k, v = "patch_embedding.proj.weight", v.reshape(v.shape[0], -1)
                vis[k] = v
        vis = fuse_qkv_and_gate_up(vis, vc.num_hidden_layers, fuse_mlp=False)
        vis = fuse_qkv_and_gate_up(vis, vc.num_global_layers, prefix="global_layers.", fuse_mlp=False)
        out.update({cls.vision_prefix + k: v for k, v in vis.items()})
        return out

    def encode_images(self, pixel_values, aspect_ratio_ids=None, aspect_ratio_mask=None, **kw):
        return self.vision_encoder_model(pixel_values, aspect_ratio_ids, aspect_ratio_mask)

    def forward(self, input_ids, attention_mask=None, position_ids=None, seq_ids=None, sampling_params=None, pixel_values=None,
                aspect_ratio_ids=None, aspect_ratio_mask=None, cross_attention_mask=None, cross_attention_states=None, **kw):
        if pixel_values is not None and cross_attention_states is None:
            cross_attention_states = self.encode_images(pixel_values, aspect_ratio_ids, aspect_ratio_mask)
        if cross_attention_mask is not None and cross_attention_mask.dim() == 4:
            # [B, T, media, tiles] -> per vision token [B, T, media*tiles*patches]
            B, T = cross_attention_mask.shape[:2]
            npatch = self.vision_model.num_patches
            cross_attention_mask = cross_attention_mask.repeat_interleave(npatch, dim=3).reshape(B, T, -1).bool()
        return NeuronBaseForImageToText.forward(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params,
                                                cross_attention_states=cross_attention_states,
                                                cross_attention_mask=cross_attention_mask, **kw)
