"""wav2vec 2.0 audio frame classifier (reference contrib/models/LaughterSegmentation: ``Wav2Vec2ForAudioFrameClassification`` on
``wav2vec2-large-xlsr-53``): raw 16 kHz waveform -> strided Conv1d feature extractor (320x down-sampling) -> projection -> grouped
convolutional position embedding -> Transformer encoder -> per-frame class logits.  Both published layouts are covered: "stable
layer norm" (large / XLSR: LayerNorm after every conv, pre-LN encoder) and base (GroupNorm after the first conv, post-LN encoder).

First user of ``NeuronEncoderApplication`` (models/encoder_base.py): one encoder sub-model, no KV cache; batches are padded along the
batch axis to the configured bucket.  The encoder GEMMs run through ``ops.linear`` (tcgen05 path on the GPU)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...config import InferenceConfig
from ...models.encoder_base import NeuronEncoderApplication, NeuronEncoderBase
from ...modules.vision import VisionAttention


class Wav2Vec2InferenceConfig(InferenceConfig):
    def get_required_attributes(self):
        return ["hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "conv_dim", "conv_stride", "conv_kernel"]

    def add_derived_config(self):
        super().add_derived_config()
        for k, d in (("feat_extract_norm", "group"), ("do_stable_layer_norm", False), ("conv_bias", False), ("layer_norm_eps", 1e-5),
                     ("num_conv_pos_embeddings", 128), ("num_conv_pos_embedding_groups", 16), ("hidden_act", "gelu"), ("num_labels", 2),
                     ("feat_extract_activation", "gelu"), ("use_weighted_layer_sum", False)):
            if getattr(self, k, None) is None:
                setattr(self, k, d)


class _EncoderLayer(nn.Module):
    def __init__(self, c, dtype, device):
        super().__init__()
        H = c.hidden_size
        self.stable = bool(c.do_stable_layer_norm)
        self.attention = VisionAttention(H, c.num_attention_heads, True, dtype, device)
        self.layer_norm = nn.LayerNorm(H, eps=c.layer_norm_eps, dtype=dtype, device=device)
        self.fc1 = nn.Linear(H, c.intermediate_size, dtype=dtype, device=device)
        self.fc2 = nn.Linear(c.intermediate_size, H, dtype=dtype, device=device)
        self.final_layer_norm = nn.LayerNorm(H, eps=c.layer_norm_eps, dtype=dtype, device=device)
        self.act = {"gelu": "gelu", "gelu_new": "gelu_tanh", "relu": "relu"}[c.hidden_act]

    def ff(self, x):
        return ops.linear(ops.activation(ops.linear(x, self.fc1.weight, self.fc1.bias), self.act), self.fc2.weight, self.fc2.bias)

    def forward(self, h):
        if self.stable:
            h = h + self.attention(self.layer_norm(h))
            return h + self.ff(self.final_layer_norm(h))
        h = self.layer_norm(h + self.attention(h))
        return self.final_layer_norm(h + self.ff(h))


class NeuronWav2Vec2FrameClassifier(NeuronEncoderBase):
    def __init__(self, config, device=None):
        super().__init__(config, device)
        c, dt = config, config.neuron_config.torch_dtype
        self.c = c
        dims = [1] + list(c.conv_dim)
        self.convs = nn.ModuleList([nn.Conv1d(dims[i], dims[i + 1], c.conv_kernel[i], c.conv_stride[i], bias=bool(c.conv_bias), dtype=dt, device=device)
                                    for i in range(len(c.conv_dim))])
        if c.feat_extract_norm == "layer":
            self.conv_norms = nn.ModuleList([nn.LayerNorm(d, dtype=dt, device=device) for d in c.conv_dim])
        else:
            self.conv_norms = nn.ModuleList([nn.GroupNorm(c.conv_dim[0], c.conv_dim[0], affine=True, dtype=dt, device=device)])
        self.proj_norm = nn.LayerNorm(c.conv_dim[-1], eps=c.layer_norm_eps, dtype=dt, device=device)
        self.projection = nn.Linear(c.conv_dim[-1], c.hidden_size, dtype=dt, device=device)
        k = c.num_conv_pos_embeddings
        self.pos_conv = nn.Conv1d(c.hidden_size, c.hidden_size, k, padding=k // 2, groups=c.num_conv_pos_embedding_groups, dtype=dt, device=device)
        self.pos_trim = 1 if k % 2 == 0 else 0
        self.encoder_norm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps, dtype=dt, device=device)
        self.layers = nn.ModuleList([_EncoderLayer(c, dt, device) for _ in range(c.num_hidden_layers)])
        self.classifier = nn.Linear(c.hidden_size, c.num_labels, dtype=dt, device=device)
        if c.use_weighted_layer_sum:
            self.layer_weights = nn.Parameter(torch.ones(c.num_hidden_layers + 1, dtype=dt, device=device) / (c.num_hidden_layers + 1))
        for p in self.parameters():
            p.requires_grad_(False)

    def num_frames(self, n_samples: int) -> int:
        for k, s in zip(self.c.conv_kernel, self.c.conv_stride):
            n_samples = (n_samples - k) // s + 1
        return n_samples

    def forward(self, input_values):
        """[B, samples] (zero-mean / unit-variance waveform) -> frame logits [B, frames, num_labels]"""
        c = self.c
        x = input_values.to(self.convs[0].weight.dtype).unsqueeze(1)
        for i, conv in enumerate(self.convs):
            x = conv(x)
            if c.feat_extract_norm == "layer":
                x = self.conv_norms[i](x.transpose(1, 2)).transpose(1, 2)
            elif i == 0:
                x = self.conv_norms[0](x)
            x = F.gelu(x)
        h = self.projection(self.proj_norm(x.transpose(1, 2)))
        pos = self.pos_conv(h.transpose(1, 2))
        if self.pos_trim:
            pos = pos[..., :-self.pos_trim]
        h = h + F.gelu(pos).transpose(1, 2)
        if not c.do_stable_layer_norm:
            h = self.encoder_norm(h)
        states = [h]
        for layer in self.layers:
            h = layer(h)
            states.append(h)
        if c.do_stable_layer_norm:
            h = self.encoder_norm(h)
            states[-1] = h
        if c.use_weighted_layer_sum:
            w = torch.softmax(self.layer_weights.float(), -1).to(h.dtype)
            h = (torch.stack(states, 1) * w.view(1, -1, 1, 1)).sum(1)
        return self.classifier(h)


class NeuronWav2Vec2ForAudioFrameClassification(NeuronEncoderApplication):
    _model_cls = NeuronWav2Vec2FrameClassifier
    _STATE_DICT_MODEL_PREFIX = "wav2vec2."
    encoder_tag = "audio_encoder_model"

    @classmethod
    def get_config_cls(cls):
        return Wav2Vec2InferenceConfig

    @staticmethod
    def load_hf_model(model_path):
        from transformers import Wav2Vec2ForAudioFrameClassification
        return Wav2Vec2ForAudioFrameClassification.from_pretrained(model_path)

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out, g, v = {}, None, None
        for k, t in sd.items():
            if k == "masked_spec_embed" or k.startswith("adapter"):
                continue
            if k.startswith("feature_extractor.conv_layers."):
                i, rest = k[len("feature_extractor.conv_layers."):].split(".", 1)
                if rest.startswith("conv."):
                    k = f"convs.{i}.{rest[5:]}"
                else:
                    j = i if config.feat_extract_norm == "layer" else "0"
                    k = f"conv_norms.{j}.{rest.split('.', 1)[1]}"
            elif k.startswith("feature_projection.layer_norm."):
                k = k.replace("feature_projection.layer_norm.", "proj_norm.")
            elif k.startswith("feature_projection.projection."):
                k = k.replace("feature_projection.", "")
            elif k.startswith("encoder.pos_conv_embed.conv."):
                rest = k[len("encoder.pos_conv_embed.conv."):]
                if rest in ("parametrizations.weight.original0", "weight_g"):
                    g = t
                    continue
                if rest in ("parametrizations.weight.original1", "weight_v"):
                    v = t
                    continue
                k = "pos_conv." + rest
            elif k.startswith("encoder.layer_norm."):
                k = k.replace("encoder.layer_norm.", "encoder_norm.")
            elif k.startswith("encoder.layers."):
                k = (k[len("encoder."):].replace(".attention.out_proj.", ".attention.o_proj.").replace(".feed_forward.intermediate_dense.", ".fc1.")
                     .replace(".feed_forward.output_dense.", ".fc2."))
            out[k] = t
        if v is not None:                    # weight norm over (out, in) per kernel tap, folded once
            out["pos_conv.weight"] = (g.float() * v.float() / v.float().norm(dim=(0, 1), keepdim=True)).to(v.dtype)
        n = config.num_hidden_layers
        from ...models.state_dict_utils import fuse_qkv_and_gate_up
        return fuse_qkv_and_gate_up(out, n, attn="attention", fuse_mlp=False)

    def forward(self, input_values, **kw):
        return self.encoder_model(input_values)
