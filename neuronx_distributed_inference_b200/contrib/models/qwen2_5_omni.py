"""Qwen2.5-Omni *thinker* with audio input (reference contrib/models/Qwen2.5-Omni-7B validates the text backbone only; the text-only
port is ``backbone_ports.NeuronQwen2_5OmniForCausalLM``).  This application adds the AUDIO tower: log-mel features are cut into
windows of ``2 * n_window`` frames, each window goes through two convolutions (stride 2) and a Whisper-style pre-LN encoder on its
own (block-diagonal attention == independent sequences, so windows are simply batched), the per-audio outputs are average-pooled by
2, normalised and projected to the text width; they replace the ``<|AUDIO|>`` placeholder tokens.

Positions: for text + audio prompts the thinker's 3-axis M-RoPE assigns the same running index to all three axes, i.e. ordinary RoPE —
the Qwen2 decoder runs unchanged.  (Image / video inputs need the Qwen2.5-VL tower and real 3-axis positions: see ``qwen2_5_vl.py``;
the talker / token2wav speech generator is out of scope.)"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...models.image_to_text_model_base import ImageToTextInferenceConfig, NeuronBaseForImageToText
from ...models.qwen2.modeling_qwen2 import NeuronQwen2ForCausalLM, NeuronQwen2Model
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from ...modules.vision import VisionAttention


class Qwen2_5OmniThinkerInferenceConfig(ImageToTextInferenceConfig):
    def get_required_attributes(self):
        return ["text_config", "audio_config"]

    def load_config(self):
        pass

    def add_derived_config(self):
        tk = getattr(self, "thinker_config", None)                 # full Omni checkpoints nest everything under thinker_config
        if tk is not None:
            for k, v in (tk.items() if isinstance(tk, dict) else vars(tk).items()):
                if k in ("text_config", "audio_config", "vision_config") and isinstance(v, dict):
                    ns = ImageToTextInferenceConfig.__new__(ImageToTextInferenceConfig)
                    for kk, vv in v.items():
                        object.__setattr__(ns, kk, vv)
                    v = ns
                if k != "neuron_config":
                    object.__setattr__(self, k, v)
        if not hasattr(self, "vision_config") or self.vision_config is None:
            object.__setattr__(self, "vision_config", ImageToTextInferenceConfig.__new__(ImageToTextInferenceConfig))
        super().add_derived_config()
        object.__setattr__(self.audio_config, "neuron_config", self.neuron_config)
        if getattr(self, "audio_token_id", None) is None:
            self.audio_token_id = getattr(self, "audio_token_index", None)


class _AudioLayer(nn.Module):
    def __init__(self, c, dtype, device):
        super().__init__()
        E = c.d_model
        self.self_attn = VisionAttention(E, c.encoder_attention_heads, True, dtype, device)
        self.self_attn_layer_norm = nn.LayerNorm(E, dtype=dtype, device=device)
        self.final_layer_norm = nn.LayerNorm(E, dtype=dtype, device=device)
        self.fc1 = nn.Linear(E, c.encoder_ffn_dim, dtype=dtype, device=device)
        self.fc2 = nn.Linear(c.encoder_ffn_dim, E, dtype=dtype, device=device)

    def forward(self, h, key_valid):
        h = h + self.self_attn(self.self_attn_layer_norm(h), key_valid=key_valid)
        return h + self.fc2(F.gelu(self.fc1(self.final_layer_norm(h))))


class NeuronQwen2_5OmniAudioEncoder(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        c = config.audio_config
        dt = config.neuron_config.torch_dtype
        self.c, self.n_window = c, int(getattr(c, "n_window", 100))
        E = c.d_model
        self.conv1 = nn.Conv1d(c.num_mel_bins, E, 3, padding=1, dtype=dt, device=device)
        self.conv2 = nn.Conv1d(E, E, 3, stride=2, padding=1, dtype=dt, device=device)
        L = int(getattr(c, "max_source_positions", 1500))
        inv = torch.exp(-(math.log(10000.0) / (E // 2 - 1)) * torch.arange(E // 2).float())
        t = torch.arange(L).float()[:, None] * inv[None, :]
        self.register_buffer("positional_embedding", torch.cat([t.sin(), t.cos()], 1).to(dt).to(device), persistent=False)
        self.layers = nn.ModuleList([_AudioLayer(c, dt, device) for _ in range(c.encoder_layers)])
        self.ln_post = nn.LayerNorm(E, dtype=dt, device=device)
        self.proj = nn.Linear(E, c.output_dim, dtype=dt, device=device)
        for p in self.parameters():
            p.requires_grad_(False)

    @staticmethod
    def output_lengths(feature_lens: torch.Tensor):
        """mel frames -> (tokens after the stride-2 conv, tokens after the pooling)"""
        after = (feature_lens - 1) // 2 + 1
        return after, (after - 2) // 2 + 1

    def forward(self, input_features, feature_attention_mask=None):
        """input_features [B, mel, T] (+ mask [B, T]) -> audio embeddings [sum(tokens), H_text] in batch order"""
        B, M, T = input_features.shape
        dev = input_features.device
        lens = feature_attention_mask.sum(1).long() if feature_attention_mask is not None else torch.full((B,), T, device=dev)
        W = 2 * self.n_window
        chunks, clen, owner = [], [], []
        for b in range(B):
            n = int(lens[b])
            for s in range(0, n, W):
                piece = input_features[b, :, s:min(s + W, n)]
                clen.append(piece.shape[1])
                chunks.append(F.pad(piece, (0, W - piece.shape[1])))
                owner.append(b)
        x = torch.stack(chunks).to(self.conv1.weight.dtype)                                 # [C, mel, W]
        clen = torch.tensor(clen, device=dev)
        m_in = (torch.arange(W, device=dev).view(1, -1) < clen.view(-1, 1))
        x = F.gelu(self.conv1(x)) * m_in.unsqueeze(1).to(x.dtype)
        x = F.gelu(self.conv2(x)).transpose(1, 2)                                            # [C, W/2, E]
        n_tok = (clen - 1) // 2 + 1
        valid = torch.arange(x.shape[1], device=dev).view(1, -1) < n_tok.view(-1, 1)
        h = x + self.positional_embedding[: x.shape[1]].unsqueeze(0)
        for layer in self.layers:
            h = layer(h, valid)
        out = []
        own = torch.tensor(owner, device=dev)
        for b in range(B):
            tok = h[own == b][valid[own == b]]                                               # [tokens of audio b, E]
            tok = F.avg_pool1d(tok.t().unsqueeze(0), 2, 2)[0].t()
            out.append(self.proj(self.ln_post(tok)))
        return torch.cat(out, 0)


class NeuronQwen2_5OmniThinkerForCausalLM(NeuronBaseForImageToText):
    _model_cls = NeuronQwen2Model
    _vision_cls = NeuronQwen2_5OmniAudioEncoder
    text_prefix = "model."
    vision_prefix = "audio_tower."
    vision_kwargs = ("feature_attention_mask",)

    @classmethod
    def get_config_cls(cls):
        return Qwen2_5OmniThinkerInferenceConfig

    @classmethod
    def _strip(cls, k):
        return k[len("thinker."):] if k.startswith("thinker.") else k

    @classmethod
    def get_state_dict(cls, path, config):
        from ...modules.checkpoint import load_state_dict
        sd = {cls._strip(k): v for k, v in load_state_dict(path).items()}
        text = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
        if "lm_head.weight" in sd:
            text["lm_head.weight"] = sd["lm_head.weight"]
        text = NeuronQwen2ForCausalLM.convert_hf_to_neuron_state_dict(text, config.get_text_config())
        if "lm_head.weight" not in text:
            text["lm_head.weight"] = text["embed_tokens.weight"].clone()
        out = {cls.text_prefix + k: v for k, v in text.items()}
        E = config.audio_config.d_model
        aud = {}
        for k, v in sd.items():
            if not k.startswith(cls.vision_prefix):
                continue
            k = k[len(cls.vision_prefix):]
            if k.startswith("audio_bos_eos_token") or "positional_embedding" in k:
                continue
            k = k.replace(".self_attn.out_proj.", ".self_attn.o_proj.")
            aud[k] = v
        for i in range(config.audio_config.encoder_layers):                                 # k_proj has no bias
            aud.setdefault(f"layers.{i}.self_attn.k_proj.bias", torch.zeros(E, dtype=aud[f"layers.{i}.self_attn.q_proj.bias"].dtype))
        aud = fuse_qkv_and_gate_up(aud, config.audio_config.encoder_layers, fuse_mlp=False)
        out.update({cls.vision_prefix + k: v for k, v in aud.items()})
        return out

    def image_token_ids(self):
        return [self.config.audio_token_id]

    def encode_images(self, input_features, feature_attention_mask=None, **kw):
        return self.vision_encoder_model(input_features, feature_attention_mask=feature_attention_mask)

    def forward(self, input_ids, attention_mask=None, position_ids=None, seq_ids=None, sampling_params=None, input_features=None,
                feature_attention_mask=None, **kw):
        return super().forward(input_ids, attention_mask, position_ids, seq_ids, sampling_params, pixel_values=input_features,
                               feature_attention_mask=feature_attention_mask, **kw)
