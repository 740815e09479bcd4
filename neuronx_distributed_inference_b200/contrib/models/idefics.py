"""IDEFICS (v1, Flamingo-style; reference contrib/models/idefics-9b-instruct): a Llama decoder with a gated cross-attention block in
front of every ``cross_layer_interval``-th layer (optional per-head q/k RMSNorm in the CROSS attention only).  A block is ``h += tanh(a1) * xattn(norm(h), image tokens)`` (zeroed for text tokens
that attend to no image) followed by ``h += tanh(a2) * mlp(norm(h))``.  Vision side: CLIP ViT (class token kept, hidden state BEFORE
the post-layernorm), optional Perceiver resampler.  Extra tokens live in "decoupled" embedding / head rows that are simply concatenated
to the base tables at load.

Serving design: like Mllama, the projected image K/V of every cross layer and the image-visibility row of the LAST prompt token are
kept per cache line (``MultimodalKVCacheManager``), so decode steps need no image inputs."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops
from ...models.image_to_text_model_base import ImageToTextInferenceConfig, NeuronBaseForImageToText
from ...models.llama.modeling_llama import NeuronLlamaMLP, NeuronLlamaModel
from ...models.model_base import DecoderLayer
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from ...modules.kvcache.multimodal_kv_cache_manager import MultimodalKVCacheManager
from ...modules.norm import RMSNorm
from ...modules.vision import PatchEmbed
from ...parallel.layers import ColumnParallelLinear, RowParallelLinear
from .llava import ClipVisionLayer


class IdeficsInferenceConfig(ImageToTextInferenceConfig):
    """IDEFICS keeps the text hyper-parameters at the top level of config.json (no ``text_config``)."""

    def get_required_attributes(self):
        return ["hidden_size", "num_attention_heads", "num_hidden_layers", "vocab_size", "vision_config"]

    def add_derived_config(self):
        super().add_derived_config()
        if getattr(self, "num_key_value_heads", None) is None:
            self.num_key_value_heads = self.num_attention_heads
        if getattr(self, "head_dim", None) is None:
            self.head_dim = self.hidden_size // self.num_attention_heads
        for k, d in (("hidden_act", "silu"), ("rms_norm_eps", 1e-6), ("max_position_embeddings", 2048), ("cross_layer_interval", 1),
                     ("additional_vocab_size", 0), ("qk_layer_norms", False), ("num_cores_per_group", 1)):
            if getattr(self, k, None) is None:
                setattr(self, k, d)
        self.base_vocab_size = self.vocab_size
        self.vocab_size = self.vocab_size + self.additional_vocab_size          # one table for base + additional tokens
        vc = self.vision_config
        if isinstance(vc, dict):
            ns = ImageToTextInferenceConfig.__new__(ImageToTextInferenceConfig)
            for k, v in vc.items():
                object.__setattr__(ns, k, v)
            object.__setattr__(ns, "neuron_config", self.neuron_config)
            object.__setattr__(self, "vision_config", ns)

    def get_text_config(self):
        return self


class IdeficsGatedCrossAttention(nn.Module):
    def __init__(self, config, idx: int, device=None):
        super().__init__()
        nc = config.neuron_config
        dt, H, D = nc.torch_dtype, config.hidden_size, config.head_dim
        self.idx, self.nh, self.D = idx, config.num_attention_heads, D
        vdim = getattr(config.vision_config, "embed_dim", H)
        self.q_proj = ColumnParallelLinear(H, self.nh * D, bias=False, gather_output=False, dtype=dt, device=device)
        self.k_proj = ColumnParallelLinear(vdim, self.nh * D, bias=False, gather_output=False, dtype=dt, device=device)
        self.v_proj = ColumnParallelLinear(vdim, self.nh * D, bias=False, gather_output=False, dtype=dt, device=device)
        self.o_proj = RowParallelLinear(self.nh * D, H, bias=False, input_is_parallel=True, dtype=dt, device=device)
        self.nh_local = self.nh // self.q_proj.tensor_parallel_group.size
        self.qk_norm = bool(config.qk_layer_norms)
        if self.qk_norm:
            self.q_layer_norm = RMSNorm(D, config.rms_norm_eps, dt, device=device)
            self.k_layer_norm = RMSNorm(D, config.rms_norm_eps, dt, device=device)
        self.input_layernorm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.post_attention_layernorm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.mlp = NeuronLlamaMLP(config, device=device)
        vec = getattr(config, "alpha_type", "float") == "vector"
        self.alpha_cross_attn = nn.Parameter(torch.zeros(H if vec else 1, dtype=dt, device=device), requires_grad=False)
        self.alpha_dense = nn.Parameter(torch.zeros(H if vec else 1, dtype=dt, device=device), requires_grad=False)

    def forward(self, h, meta, kv_mgr):
        B, T, _ = h.shape
        nh, D = self.nh_local, self.D
        lines = kv_mgr.lines_for(meta.seq_ids)
        img = meta.extras.get("image_hidden_states")
        if img is None and not kv_mgr.has_vision(self.idx):
            raise ValueError("IDEFICS needs image_hidden_states at prefill (use a zero image with an all-zero image_attention_mask for "
                             "text-only prompts, as the Hugging Face processor does)")
        q = self.q_proj(self.input_layernorm(h)).view(B, T, nh, D)
        if img is not None:
            N = img.shape[1]
            k = self.k_proj(img.to(h.dtype)).view(B, N, nh, D)
            v = self.v_proj(img.to(h.dtype)).view(B, N, nh, D).transpose(1, 2)
            if self.qk_norm:
                k = self.k_layer_norm(k)
            k = k.transpose(1, 2)
            vis = meta.extras["image_token_mask"].bool()                          # [B, T0, N]: text token x image token visibility
            if vis.shape[1] < T:
                vis = torch.cat([vis, vis.new_zeros(B, T - vis.shape[1], N)], 1)
            last = (meta.key_valid.long().sum(-1).clamp_min(1) - 1) if meta.key_valid is not None else torch.full((B,), T - 1, device=h.device)
            kv_mgr.update_vision(self.idx, lines, k, v, vis[torch.arange(B, device=h.device), last])
        else:
            k, v, row = kv_mgr.get_vision(self.idx, lines)
            vis = row.unsqueeze(1).expand(B, T, -1)
        if self.qk_norm:
            q = self.q_layer_norm(q)
        gate = vis.any(-1, keepdim=True)                                           # tokens that see at least one image
        mask = (vis | ~gate).unsqueeze(1)                                           # fully masked rows: uniform, then zeroed by the gate
        o = ops.ref.attention_with_mask(q.transpose(1, 2), k, v, mask, 1.0 / math.sqrt(D)).transpose(1, 2).reshape(B, T, nh * D)
        a = self.o_proj(o) * gate.to(h.dtype)
        h = h + torch.tanh(self.alpha_cross_attn) * a
        return h + torch.tanh(self.alpha_dense) * self.mlp(self.post_attention_layernorm(h))


class IdeficsBlock(nn.Module):
    """[gated cross-attention block] + Llama decoder layer."""
    mlp_is_moe = False

    def __init__(self, cross, decoder: DecoderLayer):
        super().__init__()
        self.cross_attn_block = cross
        self.decoder = decoder
        self.layer_idx = decoder.layer_idx

    @property
    def self_attn(self):                                # the engine sizes the KV cache from ``layers[0].self_attn``
        return self.decoder.self_attn

    def forward(self, h, meta, kv_mgr, lora=None):
        if self.cross_attn_block is not None:
            h = self.cross_attn_block(h, meta, kv_mgr)
        return self.decoder(h, meta, kv_mgr, lora)


class NeuronIdeficsTextModel(NeuronLlamaModel):
    meta_extra_keys = ("image_hidden_states", "image_token_mask")
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        dec = super().make_layer(config, i, rotary, device)
        cross = IdeficsGatedCrossAttention(config, i, device) if i % config.cross_layer_interval == 0 else None
        return IdeficsBlock(cross, dec)

    def init_inference_optimization(self, config):
        super().init_inference_optimization(config)
        nc = self.neuron_config
        if nc.is_block_kv_layout or nc.attention_dp_degree > 1:
            raise NotImplementedError("IDEFICS: contiguous KV layout only (image K/V are stored per cache line)")
        base = self.kv_mgr
        self.kv_mgr = MultimodalKVCacheManager(base.num_layers, base.num_kv_heads, base.head_dim, base.max_len, base.num_lines,
                                               nc.attention_dtype or nc.torch_dtype, self.device_,
                                               cross_attention_layers=[i for i, l in enumerate(self.layers) if l.cross_attn_block is not None])


class PerceiverBlock(nn.Module):
    """Flamingo resampler block: latents attend to [context ; latents] (LayerNorm on both, optional LayerNorm per head on q / k),
    then a bias-free ReLU MLP; both residual."""

    def __init__(self, dim, n_heads, head_dim, qk_norm, inter, dtype, device):
        super().__init__()
        self.nh, self.hd, self.qk_norm = n_heads, head_dim, qk_norm
        ln = lambda d: nn.LayerNorm(d, dtype=dtype, device=device)                                   # noqa: E731
        lin = lambda i, o: nn.Linear(i, o, bias=False, dtype=dtype, device=device)                   # noqa: E731
        self.context_layer_norm, self.latents_layer_norm = ln(dim), ln(dim)
        if qk_norm:
            self.q_layer_norm, self.k_layer_norm = ln(head_dim), ln(head_dim)
        self.q_proj, self.k_proj, self.v_proj = lin(dim, n_heads * head_dim), lin(dim, n_heads * head_dim), lin(dim, n_heads * head_dim)
        self.output_proj = lin(n_heads * head_dim, dim)
        self.ln, self.fc, self.c_proj = ln(dim), lin(dim, inter), lin(inter, dim)

    def forward(self, context, latents):
        B = context.shape[0]
        c, l = self.context_layer_norm(context), self.latents_layer_norm(latents)
        kv = torch.cat([c, l], 1)
        split = lambda x: x.view(B, x.shape[1], self.nh, self.hd).transpose(1, 2)                    # noqa: E731
        q, k, v = split(self.q_proj(l)), split(self.k_proj(kv)), split(self.v_proj(kv))
        if self.qk_norm:
            q, k = self.q_layer_norm(q), self.k_layer_norm(k)
        a = torch.softmax((q.float() * self.hd ** -0.5) @ k.float().transpose(-1, -2), -1).to(v.dtype) @ v
        latents = latents + self.output_proj(a.transpose(1, 2).flatten(2))
        return latents + self.c_proj(torch.relu(self.fc(self.ln(latents))))


class NeuronIdeficsVisionModel(nn.Module):
    """CLIP ViT; returns the encoder output for ALL tokens (class token first), before the post-layernorm."""

    def __init__(self, config, device=None):
        super().__init__()
        vc = config.vision_config
        dt = vc.neuron_config.torch_dtype
        for k, d in (("hidden_size", getattr(vc, "embed_dim", None)), ("num_channels", 3), ("hidden_act", "gelu"), ("layer_norm_eps", 1e-5)):
            if getattr(vc, k, None) is None:
                object.__setattr__(vc, k, d)
        self.vc = vc
        n = (vc.image_size // vc.patch_size) ** 2
        self.patch_embedding = PatchEmbed(vc.num_channels * vc.patch_size ** 2, vc.hidden_size, False, dt, device)
        self.class_embedding = nn.Parameter(torch.zeros(vc.hidden_size, dtype=dt, device=device), requires_grad=False)
        self.position_embedding = nn.Embedding(n + 1, vc.hidden_size, dtype=dt, device=device)
        self.pre_layrnorm = nn.LayerNorm(vc.hidden_size, eps=vc.layer_norm_eps, dtype=dt, device=device)
        self.layers = nn.ModuleList([ClipVisionLayer(vc, dt, device) for _ in range(vc.num_hidden_layers)])
        pc = getattr(config, "perceiver_config", None) or {}
        g = (lambda k, d: pc.get(k, d)) if isinstance(pc, dict) else (lambda k, d: getattr(pc, k, d))
        self.use_resampler = bool(getattr(config, "use_resampler", False) or g("use_resampler", False))
        if self.use_resampler:
            E = vc.hidden_size
            self.latents = nn.Parameter(torch.zeros(g("resampler_n_latents", 64), E, dtype=dt, device=device), requires_grad=False)
            self.blocks = nn.ModuleList([PerceiverBlock(E, g("resampler_n_heads", 16), g("resampler_head_dim", 96),
                                                        bool(g("qk_layer_norms_perceiver", False)), 4 * E, dt, device)
                                         for _ in range(g("resampler_depth", 6))])
            self.resampler_norm = nn.LayerNorm(E, dtype=dt, device=device)
        for p in self.parameters():
            p.requires_grad_(False)

    def resample(self, x):
        """[B * n_images, tokens, E] -> [B * n_images, n_latents, E]"""
        lat = self.latents.unsqueeze(0).expand(x.shape[0], -1, -1)
        for blk in self.blocks:
            lat = blk(x, lat)
        return self.resampler_norm(lat)

    def forward(self, pixel_values):
        """[B, n_images, 3, H, W] -> [B, n_images * tokens, embed_dim]; tokens = 1 + patches, or ``n_latents`` with the resampler."""
        B, n = pixel_values.shape[:2]
        x = pixel_values.flatten(0, 1)
        C, H, W = x.shape[1:]
        P = self.vc.patch_size
        x = x.reshape(B * n, C, H // P, P, W // P, P).permute(0, 2, 4, 1, 3, 5).reshape(B * n, -1, C * P * P)
        x = self.patch_embedding(x)
        x = torch.cat([self.class_embedding.view(1, 1, -1).expand(B * n, 1, -1), x], 1) + self.position_embedding.weight[: x.shape[1] + 1]
        x = self.pre_layrnorm(x)
        for layer in self.layers:
            x = layer(x)
        if self.use_resampler:
            x = self.resample(x)
        return x.reshape(B, -1, x.shape[-1])


class NeuronIdeficsForCausalLM(NeuronBaseForImageToText):
    _model_cls = NeuronIdeficsTextModel
    _vision_cls = NeuronIdeficsVisionModel
    text_prefix = ""
    vision_prefix = "vision_model."
    vision_kwargs = ()

    @classmethod
    def get_config_cls(cls):
        return IdeficsInferenceConfig

    @staticmethod
    def load_hf_model(model_path):
        from transformers import AutoModelForImageTextToText
        return AutoModelForImageTextToText.from_pretrained(model_path)

    def _split_state_dict(self, sd):
        text = {k: v for k, v in sd.items() if not k.startswith(self.vision_prefix)}
        self._vision_sd = {k[len(self.vision_prefix):]: v for k, v in sd.items() if k.startswith(self.vision_prefix)}
        return text

    @classmethod
    def get_state_dict(cls, path, config):
        from ...modules.checkpoint import load_state_dict
        sd = {cls._strip(k): v for k, v in load_state_dict(path).items()}
        n = config.num_hidden_layers
        text = {}
        for k, v in sd.items():
            if k.startswith(cls.vision_prefix) or k.startswith("perceiver_resampler."):
                continue
            if k.startswith("gated_cross_attn_layers."):
                j, rest = k[len("gated_cross_attn_layers."):].split(".", 1)
                rest = rest.replace("cross_attn.", "", 1) if rest.startswith("cross_attn.") else rest
                rest = rest.replace("alpha_cross_attn", "alpha_cross_attn").replace("alpha_dense", "alpha_dense")
                k = f"layers.{int(j) * config.cross_layer_interval}.cross_attn_block.{rest}"
                if rest.startswith("alpha_"):
                    v = v.reshape(-1)
            elif k.startswith("layers."):
                i, rest = k[len("layers."):].split(".", 1)
                k = f"layers.{i}.decoder.{rest}"
            text[k] = v
        # fuse q/k/v and gate/up of the decoder layers and gate/up of the cross blocks' MLPs
        text = fuse_qkv_and_gate_up(text, n, attn="decoder.self_attn", mlp="decoder.mlp")
        text = fuse_qkv_and_gate_up(text, n, attn="cross_attn_block.__none__", mlp="cross_attn_block.mlp")
        for a, b in (("embed_tokens.weight", "embed_tokens.additional_embedding.weight"), ("lm_head.weight", "lm_head.additional_fc.weight")):
            if b in text:
                text[a] = torch.cat([text[a], text.pop(b)], 0)
        if "lm_head.weight" not in text:
            text["lm_head.weight"] = text["embed_tokens.weight"].clone()
        out = dict(text)
        vis = {}
        for k, v in sd.items():
            if not k.startswith(cls.vision_prefix):
                continue
            k = k[len(cls.vision_prefix):]
            k = (k.replace("embeddings.class_embedding", "class_embedding").replace("embeddings.position_embedding.", "position_embedding.")
                 .replace("embeddings.patch_embedding.weight", "patch_embedding.proj.weight").replace("encoder.layers.", "layers.")
                 .replace(".self_attn.out_proj.", ".self_attn.o_proj."))
            if k == "patch_embedding.proj.weight":
                v = v.reshape(v.shape[0], -1)
            if k.startswith("post_layernorm.") or "position_ids" in k:
                continue
            vis[k] = v
        vis = fuse_qkv_and_gate_up(vis, config.vision_config.num_hidden_layers, fuse_mlp=False)
        for k, v in sd.items():                                      # Perceiver resampler lives in the vision module here
            if k.startswith("perceiver_resampler."):
                k = k[len("perceiver_resampler."):]
                if k.startswith("blocks."):
                    i, j, rest = k.split(".", 3)[1:]
                    k = f"blocks.{i}.{rest}"                          # [attention, mlp] pairs flattened into one block
                elif k.startswith("layer_norm."):
                    k = "resampler_norm." + k[len("layer_norm."):]
                vis[k] = v
        out.update({cls.vision_prefix + k: v for k, v in vis.items()})
        return out

    def encode_images(self, pixel_values, **kw):
        return self.vision_encoder_model(pixel_values)

    def forward(self, input_ids, attention_mask=None, position_ids=None, seq_ids=None, sampling_params=None, pixel_values=None,
                image_encoder_embeddings=None, perceiver_embeddings=None, image_attention_mask=None, **kw):
        """``image_attention_mask`` [B, T, n_images]: which images each text token may look at (the processor's output)."""
        if input_ids.shape[-1] > 1 and (pixel_values is not None or image_encoder_embeddings is not None or perceiver_embeddings is not None):
            emb = perceiver_embeddings if perceiver_embeddings is not None else image_encoder_embeddings
            if emb is not None:
                B, n, L, E = emb.shape
                img = emb.reshape(B, n * L, E)
            else:
                n = pixel_values.shape[1]
                img = self.encode_images(pixel_values.to(self.device))
                L = img.shape[1] // n
            if image_attention_mask is None:
                image_attention_mask = torch.ones(input_ids.shape[0], input_ids.shape[1], n, dtype=torch.bool)
            kw["image_hidden_states"] = img
            kw["image_token_mask"] = image_attention_mask.bool().unsqueeze(-1).expand(-1, -1, -1, L).reshape(*image_attention_mask.shape[:2], n * L)
        from ...models.application_base import NeuronBaseForCausalLM
        return NeuronBaseForCausalLM.forward(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params, **kw)
