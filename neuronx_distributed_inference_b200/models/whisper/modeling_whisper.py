"""Whisper speech-to-text: audio encoder application + autoregressive decoder with cross-attention.

reference: models/whisper/modeling_whisper.py:1-719 (+ utils): ``NeuronApplicationWhisper`` wraps an encoder application and
a decoder application with prefill / decode wrappers, subclassing openai-whisper's ``Whisper`` (:571-719).  Here the checkpoint
format is the Hugging Face one (``WhisperForConditionalGeneration``); the decoder re-uses the engine's decoder machinery —
contiguous KV cache addressed by ``seq_ids``, CTE/TKG runners, on-device sampling — and keeps the projected encoder K/V of
every layer in per-line buffers (computed once per utterance at prefill)."""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn

from ... import ops
from ...config import InferenceConfig, NeuronConfig
from ...modules.attention import AttentionBase
from ...modules.gqa import GroupQueryAttention_O, GroupQueryAttention_QKV
from ...modules.vision import VisionAttention, VisionMLP
from ...parallel.layers import ColumnParallelLinear, ParallelEmbedding
from ..application_base import NeuronBaseForCausalLM
from ..encoder_base import EncoderRunner
from ..model_base import NeuronBaseModel
from ..state_dict_utils import fuse_qkv_and_gate_up


class WhisperInferenceConfig(InferenceConfig):
    attribute_map = {}

    def get_required_attributes(self) -> List[str]:
        return ["d_model", "encoder_layers", "decoder_layers", "vocab_size", "num_mel_bins"]

    def add_derived_config(self):
        self.num_cores_per_group = 1
        # decoder seen through the generic decoder-model attribute names
        self.hidden_size = self.d_model
        self.num_hidden_layers = self.decoder_layers
        self.num_attention_heads = self.decoder_attention_heads
        self.num_key_value_heads = self.decoder_attention_heads
        self.head_dim = self.d_model // self.decoder_attention_heads


# ---------------------------------------------------------------------------------------------------------------------
class WhisperEncoderLayer(nn.Module):
    def __init__(self, c, dtype, device):
        super().__init__()
        self.self_attn_layer_norm = nn.LayerNorm(c.d_model, dtype=dtype, device=device)
        self.final_layer_norm = nn.LayerNorm(c.d_model, dtype=dtype, device=device)
        self.self_attn = VisionAttention(c.d_model, c.encoder_attention_heads, True, dtype, device)
        self.mlp = VisionMLP(c.d_model, c.encoder_ffn_dim, getattr(c, "activation_function", "gelu"), True, False, dtype, device)

    def forward(self, x):
        x = x + self.self_attn(self.self_attn_layer_norm(x))
        return x + self.mlp(self.final_layer_norm(x))


class NeuronWhisperEncoder(nn.Module):
    """mel ``[B, n_mels, frames]`` -> ``[B, frames/2, d_model]``."""

    def __init__(self, config, device=None):
        super().__init__()
        c, dt = config, config.neuron_config.torch_dtype
        self.conv1 = nn.Conv1d(c.num_mel_bins, c.d_model, 3, padding=1, dtype=dt, device=device)
        self.conv2 = nn.Conv1d(c.d_model, c.d_model, 3, stride=2, padding=1, dtype=dt, device=device)
        self.embed_positions = nn.Embedding(c.max_source_positions, c.d_model, dtype=dt, device=device)
        self.layers = nn.ModuleList([WhisperEncoderLayer(c, dt, device) for _ in range(c.encoder_layers)])
        self.layer_norm = nn.LayerNorm(c.d_model, dtype=dt, device=device)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, input_features):
        x = nn.functional.gelu(self.conv1(input_features.to(self.conv1.weight.dtype)))
        x = nn.functional.gelu(self.conv2(x)).transpose(1, 2)
        x = x + self.embed_positions.weight[: x.shape[1]]
        for layer in self.layers:
            x = layer(x)
        return self.layer_norm(x)


# ---------------------------------------------------------------------------------------------------------------------
class WhisperSelfAttention(AttentionBase):
    def __init__(self, c, layer_idx, device=None):
        super().__init__(c, hidden_size=c.d_model, num_attention_heads=c.decoder_attention_heads,
                         num_key_value_heads=c.decoder_attention_heads, head_dim=c.d_model // c.decoder_attention_heads,
                         rotary_emb=None, qkv_bias=True, o_bias=True, use_rope=False, layer_idx=layer_idx, device=device)


class WhisperDecoderLayer(nn.Module):
    mlp_is_moe = False

    def __init__(self, c, layer_idx, device=None):
        super().__init__()
        nc, dt = c.neuron_config, c.neuron_config.torch_dtype
        H, nh = c.d_model, c.decoder_attention_heads
        D = H // nh
        self.layer_idx, self.head_dim = layer_idx, D
        self.self_attn = WhisperSelfAttention(c, layer_idx, device)
        self.self_attn_layer_norm = nn.LayerNorm(H, dtype=dt, device=device)
        self.encoder_attn_layer_norm = nn.LayerNorm(H, dtype=dt, device=device)
        self.final_layer_norm = nn.LayerNorm(H, dtype=dt, device=device)
        self.cross_qkv = GroupQueryAttention_QKV(H, D, nh, nh, None, dt, True, None, device)
        self.cross_o = GroupQueryAttention_O(H, D, nh, nh, None, dt, True, None, device)
        self.n_heads = self.cross_qkv.n_q
        self.mlp = VisionMLP(H, c.decoder_ffn_dim, getattr(c, "activation_function", "gelu"), True, False, dt, device)
        self.num_lines = nc.kv_cache_batch_size + nc.kv_cache_padding_size + 1
        self.k_cross = self.v_cross = None

    def forward(self, h, meta, kv_mgr, lora=None):
        B, T, _ = h.shape
        nh, D = self.n_heads, self.head_dim
        h = self.self_attn(self.self_attn_layer_norm(h), meta, kv_mgr, residual=h)
        x = self.encoder_attn_layer_norm(h)
        w, b = self.cross_qkv.weight, self.cross_qkv.bias
        q = ops.linear(x, w[: nh * D], b[: nh * D]).view(B, T, nh, D)
        enc = meta.extras.get("encoder_hidden_states")
        lines = kv_mgr.lines_for(meta.seq_ids).long().clamp(0, self.num_lines - 1)
        if enc is not None:
            S = enc.shape[1]
            kv = ops.linear(enc.to(h.dtype), w[nh * D:], b[nh * D:]).view(B, S, 2 * nh, D)
            k, v = kv[:, :, :nh].transpose(1, 2), kv[:, :, nh:].transpose(1, 2)
            if self.k_cross is None or self.k_cross.shape[2] != S:
                self.k_cross = k.new_zeros(self.num_lines, nh, S, D)
                self.v_cross = v.new_zeros(self.num_lines, nh, S, D)
            self.k_cross[lines], self.v_cross[lines] = k, v
        else:
            k, v = self.k_cross[lines], self.v_cross[lines]
        mask = torch.ones(1, 1, 1, k.shape[2], dtype=torch.bool, device=h.device)
        o = ops.ref.attention_with_mask(q.transpose(1, 2), k, v, mask, 1.0 / math.sqrt(D))
        h = self.cross_o(o.transpose(1, 2).reshape(B, T, nh * D), h)
        return h + self.mlp(self.final_layer_norm(h))


class NeuronWhisperDecoderModel(NeuronBaseModel):
    meta_extra_keys = ("encoder_hidden_states",)
    graph_safe = False

    def setup_attr_for_model(self, config):
        nc = config.neuron_config
        self.tp_degree, self.hidden_size = nc.tp_degree, config.d_model
        self.num_attention_heads = self.num_key_value_heads = config.decoder_attention_heads
        self.max_batch_size, self.buckets = nc.max_batch_size, nc.buckets

    def init_model(self, config):
        nc, dt, dev = config.neuron_config, config.neuron_config.torch_dtype, self.device_
        self.embed_tokens = ParallelEmbedding(config.vocab_size, config.d_model, getattr(config, "pad_token_id", None), dtype=dt,
                                              device=dev, shard_across_embedding=True, pad=True, tensor_model_parallel_group=self.tp_group)
        self.embed_positions = nn.Embedding(config.max_target_positions, config.d_model, dtype=dt, device=dev)
        self.embed_positions.weight.requires_grad_(False)
        self.layers = nn.ModuleList([WhisperDecoderLayer(config, i, dev) for i in range(config.decoder_layers)])
        self.norm = nn.LayerNorm(config.d_model, dtype=dt, device=dev)
        for p in self.norm.parameters():
            p.requires_grad_(False)
        self.lm_head = ColumnParallelLinear(config.d_model, config.vocab_size, bias=False, gather_output=False, dtype=dt, device=dev,
                                            pad=True, tensor_model_parallel_group=self.tp_group)

    def forward(self, input_ids, attention_mask=None, position_ids=None, *a, **kw):
        self._pos = position_ids
        return super().forward(input_ids, attention_mask, position_ids, *a, **kw)

    def embed(self, input_ids, inputs_embeds=None, vision_embeddings=None, vision_mask=None):
        h = self.embed_tokens(input_ids)
        pos = self._pos if self._pos is not None else torch.arange(input_ids.shape[1], device=input_ids.device).unsqueeze(0)
        return h + self.embed_positions(pos.long().clamp(0, self.embed_positions.num_embeddings - 1))

    def final_hidden(self, h):
        return self.norm(h)

    def compute_logits(self, h):
        return self.lm_head(self.norm(h))


class NeuronApplicationWhisper(NeuronBaseForCausalLM):
    """``forward(decoder_input_ids, ..., input_features=mel)``; ``generate(input_features, decoder_input_ids, max_new_tokens)``."""
    _model_cls = NeuronWhisperDecoderModel
    _STATE_DICT_MODEL_PREFIX = "model."

    @classmethod
    def get_config_cls(cls):
        return WhisperInferenceConfig

    @staticmethod
    def load_hf_model(model_path):
        from transformers import WhisperForConditionalGeneration
        return WhisperForConditionalGeneration.from_pretrained(model_path)

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        H = config.d_model

        def fuse(prefix, dst):
            q, k, v = (sd[f"{prefix}.{p}_proj.weight"] for p in "qkv")
            out[f"{dst}.weight"] = torch.cat([q, k, v], 0)
            out[f"{dst}.bias"] = torch.cat([sd[f"{prefix}.q_proj.bias"], torch.zeros(H, dtype=q.dtype), sd[f"{prefix}.v_proj.bias"]], 0)
        for i in range(config.encoder_layers):
            p = f"encoder.layers.{i}"
            fuse(f"{p}.self_attn", f"{p}.self_attn.qkv_proj")
            out[f"{p}.self_attn.o_proj.weight"], out[f"{p}.self_attn.o_proj.bias"] = sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"]
            for n in ("fc1", "fc2"):
                out[f"{p}.mlp.{n}.weight"], out[f"{p}.mlp.{n}.bias"] = sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"]
            for n in ("self_attn_layer_norm", "final_layer_norm"):
                out[f"{p}.{n}.weight"], out[f"{p}.{n}.bias"] = sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"]
        for k in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "embed_positions.weight", "layer_norm.weight", "layer_norm.bias"):
            out[f"encoder.{k}"] = sd[f"encoder.{k}"]
        for i in range(config.decoder_layers):
            p, d = f"decoder.layers.{i}", f"layers.{i}"
            fuse(f"{p}.self_attn", f"{d}.self_attn.qkv_proj")
            out[f"{d}.self_attn.o_proj.weight"], out[f"{d}.self_attn.o_proj.bias"] = sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"]
            fuse(f"{p}.encoder_attn", f"{d}.cross_qkv")
            out[f"{d}.cross_o.weight"], out[f"{d}.cross_o.bias"] = sd[f"{p}.encoder_attn.out_proj.weight"], sd[f"{p}.encoder_attn.out_proj.bias"]
            for n in ("fc1", "fc2"):
                out[f"{d}.mlp.{n}.weight"], out[f"{d}.mlp.{n}.bias"] = sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"]
            for n in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"):
                out[f"{d}.{n}.weight"], out[f"{d}.{n}.bias"] = sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"]
        out["embed_tokens.weight"] = sd["decoder.embed_tokens.weight"]
        out["embed_positions.weight"] = sd["decoder.embed_positions.weight"]
        out["norm.weight"], out["norm.bias"] = sd["decoder.layer_norm.weight"], sd["decoder.layer_norm.bias"]
        out["lm_head.weight"] = sd.get("proj_out.weight", sd["decoder.embed_tokens.weight"])
        return out

    def checkpoint_loader_fn(self, mmap: bool = False) -> dict:
        sd = super().checkpoint_loader_fn(mmap)
        self._encoder_sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
        return {k: v for k, v in sd.items() if not k.startswith("encoder.")}

    def _post_load(self, model):
        super()._post_load(model)
        from ...modules.checkpoint import load_sharded
        with torch.device(self.device):
            self.encoder = NeuronWhisperEncoder(self.config, self.device).eval()
        esd = getattr(self, "_encoder_sd", None)
        if esd:
            load_sharded(self.encoder, esd, self.neuron_config.torch_dtype, strict=False)
            self._encoder_sd = None
        else:
            main, self.model = self.model, self.encoder
            try:
                self.init_random_weights(7)
            finally:
                self.model = main

    def _build_runners(self):
        super()._build_runners()
        self.encoder_model = EncoderRunner("encoder_model", self.encoder, None, 0, self.device)
        self.models.append(self.encoder_model)

    def encode(self, input_features):
        return self.encoder_model(input_features)

    def forward(self, input_ids, attention_mask=None, position_ids=None, seq_ids=None, sampling_params=None, input_features=None,
                encoder_hidden_states=None, **kw):
        if input_features is not None and encoder_hidden_states is None:
            encoder_hidden_states = self.encode(input_features)
        return super().forward(input_ids, attention_mask, position_ids, seq_ids, sampling_params,
                               encoder_hidden_states=encoder_hidden_states, **kw)

    @torch.no_grad()
    def generate(self, input_features, decoder_input_ids=None, max_new_tokens: int = 32, eos_token_id: Optional[int] = None):
        """Greedy transcription loop -> token ids ``[B, prompt + new]`` (pads with eos after the end)."""
        B = input_features.shape[0]
        eos = eos_token_id if eos_token_id is not None else getattr(self.config, "eos_token_id", None)
        if decoder_input_ids is None:
            decoder_input_ids = torch.full((B, 1), getattr(self.config, "decoder_start_token_id", 0), dtype=torch.long)
        self.reset()
        seq = decoder_input_ids.clone()
        out = self(seq, input_features=input_features)
        done = torch.zeros(B, dtype=torch.bool)
        for step in range(max_new_tokens):
            nxt = (out.tokens.reshape(B, -1)[:, -1] if out.tokens is not None else out.logits[:, -1].argmax(-1)).cpu().long()
            if eos is not None:
                nxt = torch.where(done, torch.full_like(nxt, eos), nxt)
                done |= nxt == eos
            seq = torch.cat([seq, nxt.view(B, 1)], 1)
            if bool(done.all()) or step == max_new_tokens - 1:
                break
            pos = torch.full((B, 1), seq.shape[1] - 1, dtype=torch.int32)
            out = self(nxt.view(B, 1), position_ids=pos)
        return seq

    transcribe_tokens = generate
