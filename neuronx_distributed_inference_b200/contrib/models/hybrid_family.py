"""Hybrid decoders: attention layers on the engine's KV cache interleaved with fixed-state sequence mixers whose state lives in a
:class:`RecurrentStateCache` attached to the KV manager (same cache lines, same reset).

* **LFM2** — double-gated short causal convolution (``y = C * conv_L(B * x)``, depthwise, L = 3) in most layers, GQA attention with
  per-head q/k RMSNorm in the rest; SwiGLU with the auto-adjusted width.
* **RecurrentGemma (Griffin)** — two recurrent blocks (``linear_x -> causal conv4 -> RG-LRU``, gated by ``gelu(linear_y)``) per local
  (sliding-window, MQA, half-rotary) attention block; Gemma norms / embedding scale / GeGLU with biases / logit soft-cap.
  RG-LRU: ``a = exp(-8 * sigmoid(W_a x) * softplus(L))``, ``h_t = a_t h_{t-1} + sqrt(1 - a_t^2) * (sigmoid(W_i x) * x_t)`` with
  block-diagonal (per head) gate matrices; the state is fp32 and reset at position 0.
* **Bamba** — Mamba-2 layers with a few GQA attention layers (half-rotary) in between, SwiGLU after every mixer; the mixer gates first
  and then applies one RMSNorm over the whole inner width.
* **Granite-4.0 hybrid (``granitemoehybrid``)** — Bamba-style Mamba-2 / attention stack (RoPE or no positions at all) with the Granite
  multipliers and, per layer, an always-on shared SwiGLU plus an optional top-k MoE.
* **Mamba / Falcon-Mamba** — attention-free Mamba-1 (per-channel selective SSM; Falcon-Mamba adds weight-free RMS on B, C, dt).
* **Jamba** — Mamba-1 layers (learned norms on dt / B / C) with a position-free attention layer every few blocks and a top-k MoE
  SwiGLU on alternating layers.
* **Mamba-2 (Codestral-Mamba)** — attention-free: ``h += mixer(norm(h))`` per layer; the engine keeps a token-sized dummy KV cache.
* **Falcon-H1** — every layer runs a Mamba-2 mixer and GQA attention IN PARALLEL on the same normed input and sums them; muP
  multipliers everywhere (all linear, folded into the weights at load).
* **Nemotron-H / Nemotron-3-Nano (``nemotron_h``)** — ONE mixer per block (Mamba-2 with a grouped gated norm, GQA attention without
  positions, squared-ReLU MLP, or a sigmoid-routed MoE of non-gated squared-ReLU experts + a shared expert).  Mamba-2: ``h_t = exp(dt_t A) h_{t-1} + dt_t B_t x_t``,
  ``y_t = C_t h_t + D x_t`` per head with grouped B/C, causal conv4 + SiLU in front, gated (grouped) RMSNorm or SiLU gate behind.
reference ports: contrib/models/{lfm2-2.6b, recurrentgemma-2b-it, Falcon-H1-0.5B-Instruct}/src."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...models.llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel
from ...models.model_base import DecoderLayer
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from ...modules.kvcache.recurrent_state_cache import RecurrentStateCache
from ...modules.mlp import GatedMLP
from ...modules.norm import RMSNorm
from ...parallel.layers import ColumnParallelLinear, RowParallelLinear


def _last_valid(meta, B, T, device):
    """Number of valid (non-padding) tokens per row of a right-padded prefill."""
    if getattr(meta, "key_valid", None) is not None:
        return meta.key_valid.long().sum(-1).clamp(min=1)
    return torch.full((B,), T, dtype=torch.long, device=device)


class _HybridModel(NeuronLlamaModel):
    """Adds the recurrent-state cache.  Layers declare ``state_specs() -> {name: per-line shape}``."""
    graph_safe = False

    def kv_heads_per_rank(self):
        return next((l.self_attn.n_kv for l in self.layers if hasattr(l, "self_attn")), 1)      # attention-free stacks keep a token KV cache

    def kv_head_dim(self):
        return next((l.self_attn.head_dim for l in self.layers if hasattr(l, "self_attn")), 8)

    def init_inference_optimization(self, config):
        super().init_inference_optimization(config)
        nc = self.neuron_config
        if nc.is_block_kv_layout or nc.speculation_length or nc.is_medusa or nc.is_chunked_prefill or nc.attention_dp_degree > 1:
            raise NotImplementedError("hybrid recurrent layers: contiguous KV cache, one token per decode step")
        if nc.padding_side != "right":
            raise NotImplementedError("hybrid recurrent layers take right-padded prompts (the state after the LAST VALID token is kept)")
        specs = {}
        for layer in self.layers:
            specs.update(getattr(layer, "state_specs", dict)())
        self.kv_mgr.states = RecurrentStateCache(specs, self.kv_mgr.num_lines, nc.torch_dtype, self.device_)


# ---------------------------------------------------------------------------------------------------------------------- LFM2
class Lfm2InferenceConfig(LlamaInferenceConfig):
    def get_required_attributes(self):
        return ["hidden_size", "num_attention_heads", "num_hidden_layers", "num_key_value_heads", "vocab_size", "layer_types"]

    def add_derived_config(self):
        self.rms_norm_eps = getattr(self, "norm_eps", 1e-5)
        self.hidden_act = "silu"
        I = self.intermediate_size
        if getattr(self, "block_auto_adjust_ff_dim", True):
            I = int(2 * I / 3)
            mult = getattr(self, "block_ffn_dim_multiplier", None)
            if mult is not None:
                I = int(mult * I)
                m = getattr(self, "block_multiple_of", 256)
                I = m * ((I + m - 1) // m)
        self.intermediate_size = I
        self.block_auto_adjust_ff_dim = False         # idempotent across save / load of the derived config
        super().add_derived_config()


class Lfm2ShortConvLayer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, device=None):
        super().__init__()
        nc = config.neuron_config
        dt, H, self.L = nc.torch_dtype, config.hidden_size, int(getattr(config, "conv_L_cache", 3))
        bias = bool(getattr(config, "conv_bias", False))
        self.in_proj = ColumnParallelLinear(H, 3 * H, bias=bias, gather_output=False, dtype=dt, device=device, stride=3)
        self.out_proj = RowParallelLinear(H, H, bias=bias, input_is_parallel=True, dtype=dt, device=device)
        tp = self.in_proj.tensor_parallel_group
        self.Hl = H // tp.size
        self.conv_weight = nn.Parameter(torch.zeros(self.Hl, self.L, dtype=dt, device=device), requires_grad=False)
        self.conv_weight.partition_dim, self.conv_weight.tp_group = 0, tp
        self.conv_bias = None
        if bias:
            self.conv_bias = nn.Parameter(torch.zeros(self.Hl, dtype=dt, device=device), requires_grad=False)
            self.conv_bias.partition_dim, self.conv_bias.tp_group = 0, tp
        self.mlp = GatedMLP(H, config.intermediate_size, "silu", dt, device=device)
        self.operator_norm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.ffn_norm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.layer_idx, self.state = i, f"conv{i}"

    def state_specs(self):
        return {self.state: (self.L - 1, self.Hl)}

    def forward(self, h, meta, kv_mgr, lora=None):
        B, T, _ = h.shape
        L = self.L
        Bm, C, x = self.in_proj(self.operator_norm(h)).chunk(3, dim=-1)
        Bx = Bm * x                                                            # [B, T, Hl]
        lines = kv_mgr.lines_for(meta.seq_ids)
        w = self.conv_weight.t().unsqueeze(0)                                  # [1, L, Hl]
        if meta.is_prefill:
            if meta.has_prefix:
                raise NotImplementedError("LFM2 short convolution with a cached prefix")
            pad = F.pad(Bx, (0, 0, L - 1, 0))                                  # causal: L-1 zeros in front
            conv = sum(pad[:, j:j + T] * w[:, j:j + 1] for j in range(L))
            n = _last_valid(meta, B, T, h.device)                              # state = the last L-1 VALID inputs of every row
            idx = (n.view(B, 1) + torch.arange(L - 1, device=h.device).view(1, -1)).unsqueeze(-1).expand(B, L - 1, Bx.shape[-1])
            kv_mgr.states.write(self.state, lines, pad.gather(1, idx))
        else:
            if T != 1:
                raise NotImplementedError("LFM2 short convolution takes one new token per decode step")
            win = torch.cat([kv_mgr.states.read(self.state, lines).to(Bx.dtype), Bx], 1)      # [B, L, Hl]
            conv = (win * w).sum(1, keepdim=True)
            kv_mgr.states.write(self.state, lines, win[:, 1:])
        if self.conv_bias is not None:
            conv = conv + self.conv_bias
        h = self.out_proj(C * conv, h)
        n2 = self.ffn_norm
        return self.mlp(h, norm_weight=n2.weight, norm_eps=n2.variance_epsilon, residual=h)


class _Lfm2Attention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, qk_norm="rms_pre_rope", qk_norm_eps=config.rms_norm_eps, **over)


class NeuronLfm2Model(_HybridModel):
    attention_cls = _Lfm2Attention

    def make_layer(self, config, i, rotary, device):
        if config.layer_types[i] != "full_attention":
            return Lfm2ShortConvLayer(config, i, device)
        dt = config.neuron_config.torch_dtype
        return DecoderLayer(self.attention_cls(config, i, rotary, device=device),
                            GatedMLP(config.hidden_size, config.intermediate_size, "silu", dt, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device),
                            RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device), i)


class NeuronLfm2ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronLfm2Model

    @classmethod
    def get_config_cls(cls):
        return Lfm2InferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        ren = ((".feed_forward.w1.", ".mlp.gate_proj."), (".feed_forward.w3.", ".mlp.up_proj."), (".feed_forward.w2.", ".mlp.down_proj."),
               (".self_attn.out_proj.", ".self_attn.o_proj."), (".conv.in_proj.", ".in_proj."), (".conv.out_proj.", ".out_proj."),
               ("embedding_norm.", "norm."))
        out = {}
        for k, v in sd.items():
            for a, b in ren:
                k = k.replace(a, b)
            if k.endswith(".conv.conv.weight"):
                k, v = k.replace(".conv.conv.weight", ".conv_weight"), v.squeeze(1)            # [H, 1, L] -> [H, L]
            elif k.endswith(".conv.conv.bias"):
                k = k.replace(".conv.conv.bias", ".conv_bias")
            out[k] = v
        for i, kind in enumerate(config.layer_types):
            if kind == "full_attention":                                                        # attention layers use the Llama names
                for a, b in ((f"layers.{i}.operator_norm.", f"layers.{i}.input_layernorm."),
                             (f"layers.{i}.ffn_norm.", f"layers.{i}.post_attention_layernorm.")):
                    for k in [k for k in out if k.startswith(a)]:
                        out[b + k[len(a):]] = out.pop(k)
        return fuse_qkv_and_gate_up(out, config.num_hidden_layers)


# ---------------------------------------------------------------------------------------------------------------------- LFM2-MoE
class Lfm2MoeRouter(nn.Module):
    """sigmoid affinities; the per-expert bias steers the SELECTION only; weights renormalised (+1e-6) and scaled."""

    def __init__(self, config, device=None):
        super().__init__()
        self.E, self.top_k = config.num_experts, config.num_experts_per_tok
        self.norm, self.scaling = bool(getattr(config, "norm_topk_prob", True)), float(getattr(config, "routed_scaling_factor", 1.0))
        self.linear_router = nn.Linear(config.hidden_size, self.E, bias=False, dtype=torch.float32, device=device)
        self.linear_router.weight.requires_grad_(False)
        self.use_bias = bool(getattr(config, "use_expert_bias", True))
        self.register_buffer("expert_bias", torch.zeros(self.E, dtype=torch.float32, device=device))

    def forward(self, x):
        logits = F.linear(x.float(), self.linear_router.weight)
        s = logits.sigmoid()
        idx = (s + self.expert_bias if self.use_bias else s).topk(self.top_k, -1)[1]
        w = s.gather(1, idx)
        if self.norm:
            w = w / (w.sum(-1, keepdim=True) + 1e-6)
        return logits, w * self.scaling, idx


class Lfm2MoeInferenceConfig(Lfm2InferenceConfig):
    def add_derived_config(self):
        self.block_auto_adjust_ff_dim = False                       # LFM2-MoE states its widths directly
        super().add_derived_config()

    @classmethod
    def get_neuron_config_cls(cls):
        from ...config import MoENeuronConfig
        return MoENeuronConfig


class NeuronLfm2MoeModel(NeuronLfm2Model):
    """LFM2 whose feed-forward becomes a sigmoid-routed MoE from layer ``num_dense_layers`` on."""

    def make_layer(self, config, i, rotary, device):
        from ...modules.moe import ExpertMLPs, MoE
        layer = super().make_layer(config, i, rotary, device)
        if i >= int(getattr(config, "num_dense_layers", 0)):
            layer.mlp = MoE(Lfm2MoeRouter(config, device),
                            ExpertMLPs(config.num_experts, config.hidden_size, config.moe_intermediate_size, "silu",
                                       config.neuron_config.torch_dtype, device=device))
            layer.mlp_is_moe = True
        return layer


class NeuronLfm2MoeForCausalLM(NeuronLfm2ForCausalLM):
    _model_cls = NeuronLfm2MoeModel

    @classmethod
    def get_config_cls(cls):
        return Lfm2MoeInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        from ...models.state_dict_utils import convert_moe_experts
        sd = {k.replace(".feed_forward.expert_bias", ".mlp.router.expert_bias"): v for k, v in sd.items()}
        moe_layers = range(int(getattr(config, "num_dense_layers", 0)), config.num_hidden_layers)
        sd = convert_moe_experts(sd, config.num_hidden_layers, config.num_experts, moe_prefixes=("feed_forward",), gate_names=("gate",),
                                 w_names=("w1", "w3", "w2"), layers=moe_layers)
        return NeuronLfm2ForCausalLM.convert_hf_to_neuron_state_dict(sd, config)


# ---------------------------------------------------------------------------------------------------------------------- RecurrentGemma
class RecurrentGemmaInferenceConfig(LlamaInferenceConfig):
    def get_required_attributes(self):
        return ["hidden_size", "num_attention_heads", "num_hidden_layers", "vocab_size", "lru_width"]

    def add_derived_config(self):
        self.hidden_act = getattr(self, "hidden_activation", "gelu_pytorch_tanh")
        if getattr(self, "num_key_value_heads", None) is None:
            self.num_key_value_heads = 1
        if not getattr(self, "layers_block_type", None):
            pat = list(getattr(self, "block_types", ("recurrent", "recurrent", "attention")))
            self.layers_block_type = [pat[i % len(pat)] for i in range(self.num_hidden_layers)]
        super().add_derived_config()


def _head_sharded(t: nn.Parameter, group):
    t.partition_dim, t.tp_group = 0, group
    return t


class GriffinRecurrentBlock(nn.Module):
    """x/y branches -> depthwise causal conv (width 4) on x -> RG-LRU -> gate by gelu(y) -> output projection."""

    def __init__(self, config, i, device=None):
        super().__init__()
        nc = config.neuron_config
        dt, H, W = nc.torch_dtype, config.hidden_size, config.lru_width
        self.linear_x = ColumnParallelLinear(H, W, bias=True, gather_output=False, dtype=dt, device=device)
        self.linear_y = ColumnParallelLinear(H, W, bias=True, gather_output=False, dtype=dt, device=device)
        self.linear_out = RowParallelLinear(W, H, bias=True, input_is_parallel=True, dtype=dt, device=device)
        g = self.linear_x.tensor_parallel_group
        nh = config.num_attention_heads
        assert nh % g.size == 0, "RG-LRU gate blocks are sharded by head"
        self.nh, self.bw, self.Wl, self.K = nh // g.size, W // nh, W // g.size, int(getattr(config, "conv1d_width", 4))
        mk = lambda *shape: _head_sharded(nn.Parameter(torch.zeros(*shape, dtype=dt, device=device), requires_grad=False), g)  # noqa: E731
        self.conv_weight, self.conv_bias = mk(self.Wl, self.K), mk(self.Wl)
        self.recurrent_param = mk(self.Wl)
        self.input_gate_weight, self.input_gate_bias = mk(self.nh, self.bw, self.bw), mk(self.nh, self.bw)
        self.recurrent_gate_weight, self.recurrent_gate_bias = mk(self.nh, self.bw, self.bw), mk(self.nh, self.bw)
        self.conv_state, self.lru_state = f"rg_conv{i}", f"rg_lru{i}"

    def state_specs(self):
        return {self.conv_state: (self.K - 1, self.Wl), self.lru_state: ((self.Wl,), torch.float32)}

    def _gates(self, x):
        B, T, _ = x.shape
        xh = x.reshape(B * T, self.nh, self.bw).transpose(0, 1).float()                     # [nh, B*T, bw]
        gi = torch.baddbmm(self.input_gate_bias.float().unsqueeze(1), xh, self.input_gate_weight.float())
        gr = torch.baddbmm(self.recurrent_gate_bias.float().unsqueeze(1), xh, self.recurrent_gate_weight.float())
        back = lambda t: t.transpose(0, 1).reshape(B, T, self.Wl)                           # noqa: E731
        return torch.sigmoid(back(gi)), torch.sigmoid(back(gr))

    def forward(self, xn, meta, kv_mgr):
        """xn: normed hidden [B, T, H] -> block output before the residual [B, T, H]."""
        B, T, _ = xn.shape
        K = self.K
        y = ops.activation(self.linear_y(xn), "gelu_tanh")
        x = self.linear_x(xn)                                                               # [B, T, Wl]
        lines = kv_mgr.lines_for(meta.seq_ids)
        states = kv_mgr.states
        w = self.conv_weight.t().unsqueeze(0)                                               # [1, K, Wl]
        pos = meta.position_ids.long()
        if meta.is_prefill:
            if meta.has_prefix:
                raise NotImplementedError("RG-LRU with a cached prefix")
            n = _last_valid(meta, B, T, xn.device)
            pad = F.pad(x, (0, 0, K - 1, 0))
            conv = sum(pad[:, j:j + T] * w[:, j:j + 1] for j in range(K)) + self.conv_bias
            idx = (n.view(B, 1) + torch.arange(K - 1, device=xn.device).view(1, -1)).unsqueeze(-1).expand(B, K - 1, self.Wl)
            states.write(self.conv_state, lines, pad.gather(1, idx))
            valid = (torch.arange(T, device=xn.device).view(1, T) < n.view(B, 1)).unsqueeze(-1)
            h0 = torch.zeros(B, self.Wl, dtype=torch.float32, device=xn.device)
        else:
            if T != 1:
                raise NotImplementedError("RG-LRU takes one new token per decode step")
            win = torch.cat([states.read(self.conv_state, lines).to(x.dtype), x], 1)
            conv = (win * w).sum(1, keepdim=True) + self.conv_bias
            states.write(self.conv_state, lines, win[:, 1:])
            valid = torch.ones(B, 1, 1, dtype=torch.bool, device=xn.device)
            h0 = states.read(self.lru_state, lines).float()
        gi, gr = self._gates(conv)
        log_a = -8.0 * gr * F.softplus(self.recurrent_param.float())
        a = torch.exp(log_a)
        reset = (pos == 0).unsqueeze(-1)
        mult = torch.where(reset, torch.ones_like(a), torch.sqrt((1.0 - torch.exp(2.0 * log_a)).clamp(min=0.0)))
        u = (conv * gi.to(conv.dtype) * mult.to(conv.dtype)).float()                        # HF rounds the gated input to the model dtype
        a = torch.where(reset, torch.zeros_like(a), a)
        a = torch.where(valid, a, torch.ones_like(a))                                       # padding steps carry the state through
        u = torch.where(valid, u, torch.zeros_like(u))
        hs, h = [], h0
        for t in range(T):
            h = a[:, t] * h + u[:, t]
            hs.append(h)
        states.write(self.lru_state, lines, h)
        out = torch.stack(hs, 1).to(xn.dtype)
        return self.linear_out(out * y)


class _GriffinAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, qkv_bias=bool(getattr(config, "attention_bias", False)), o_bias=True,
                         sliding_window=getattr(config, "attention_window_size", None), **over)


class GriffinLayer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        self.recurrent = config.layers_block_type[i] == "recurrent"
        if self.recurrent:
            self.temporal_block = GriffinRecurrentBlock(config, i, device)
        else:
            self.self_attn = _GriffinAttention(config, i, rotary, device=device)
        mk = lambda: RMSNorm(config.hidden_size, config.rms_norm_eps, dt, offset=1.0, device=device)   # noqa: E731
        self.temporal_pre_norm, self.channel_pre_norm = mk(), mk()
        self.mlp = GatedMLP(config.hidden_size, config.intermediate_size // 2, config.hidden_act, dt, bias=True, device=device)
        self.layer_idx = i

    def state_specs(self):
        return self.temporal_block.state_specs() if self.recurrent else {}

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.temporal_pre_norm
        if self.recurrent:
            h = h + self.temporal_block(n(h), meta, kv_mgr)
        else:
            h = self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon, norm_offset=n.offset, residual=h)
        n = self.channel_pre_norm
        return self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon, norm_offset=n.offset, residual=h)


class NeuronRecurrentGemmaModel(_HybridModel):
    def make_rotary(self, config, device):
        from ...models.llama.modeling_llama import rope_theta_of
        from ...modules.rope import RotaryEmbedding
        rp = getattr(config, "rope_parameters", None) or {}
        frac = getattr(config, "partial_rotary_factor", None) or (rp.get("partial_rotary_factor") if isinstance(rp, dict) else None) or 0.5
        return RotaryEmbedding(int(config.head_dim * float(frac)), max(config.max_position_embeddings, config.neuron_config.seq_len),
                               rope_theta_of(config), None, device=device)

    def make_layer(self, config, i, rotary, device):
        return GriffinLayer(config, i, rotary, device)

    def init_model(self, config):
        if not hasattr(config, "max_position_embeddings"):
            config.max_position_embeddings = config.neuron_config.seq_len
        super().init_model(config)
        dt = config.neuron_config.torch_dtype
        self.norm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, offset=1.0, device=self.device_)
        self.embed_scale = float(torch.tensor(config.hidden_size ** 0.5, dtype=torch.bfloat16))
        self.final_logit_softcap = getattr(config, "logits_soft_cap", None)


class NeuronRecurrentGemmaForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronRecurrentGemmaModel

    @classmethod
    def get_config_cls(cls):
        return RecurrentGemmaInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            k = k.replace("final_norm.", "norm.").replace(".mlp_block.", ".mlp.")
            if ".temporal_block." in k:
                i = int(k.split(".")[1])
                if config.layers_block_type[i] == "attention":
                    k = k.replace(".temporal_block.", ".self_attn.")
                else:
                    k = k.replace(".rg_lru.", ".")
                    if k.endswith(".conv_1d.weight"):
                        k, v = k.replace(".conv_1d.weight", ".conv_weight"), v.squeeze(1)
                    elif k.endswith(".conv_1d.bias"):
                        k = k.replace(".conv_1d.bias", ".conv_bias")
            out[k] = v
        return fuse_qkv_and_gate_up(out, config.num_hidden_layers)


# ---------------------------------------------------------------------------------------------------------------------- Falcon-H1
class Mamba2Mixer(nn.Module):
    """Mamba-2 (SSD) block in recurrent form, tensor parallel over HEADS: rank ``r`` owns ``nh / tp`` heads — their gate / x rows of
    ``in_proj``, their ``dt`` / ``A`` / ``D`` entries, the matching convolution channels and ``out_proj`` columns (one all-reduce at the
    end) — plus the B / C rows of the group(s) those heads belong to (replicated among the ranks of a group when there are fewer groups
    than ranks).  The gated RMSNorm is per group; when a group spans several ranks its sum of squares is combined across them."""

    def __init__(self, config, i, device=None, gated_norm=None, norm_before_gate=None, norm_groups=None, out_bias=None):
        """The keyword overrides cover the checkpoint families that share this block: Falcon-H1 (grouped norm after / before the gate,
        optional), Bamba / Granite-4 (gate first, ONE RMSNorm over the whole inner width)."""
        super().__init__()
        from ...parallel.state import get_tensor_model_parallel_group
        dt = config.neuron_config.torch_dtype
        H = config.hidden_size
        self.tp_group = g = get_tensor_model_parallel_group()
        tp, r = g.size, g.rank
        nh, hd, G, N = config.mamba_n_heads, config.mamba_d_head, config.mamba_n_groups, config.mamba_d_state
        I = config.mamba_d_ssm if getattr(config, "mamba_d_ssm", None) is not None else int(config.mamba_expand * H)
        assert I == nh * hd and nh % tp == 0 and nh % G == 0 and (G % tp == 0 or tp % G == 0), "Mamba-2 heads / groups vs tp degree"
        self.nh, self.hd, self.N, self.K = nh // tp, hd, N, config.mamba_d_conv
        self.G = max(G // tp, 1)                                   # groups whose B / C this rank needs
        self.I = self.nh * hd
        self.conv_dim = self.I + 2 * self.G * N
        NG = G if norm_groups is None else norm_groups             # groups of the gated RMSNorm (may differ from the B / C groups)
        assert NG % tp == 0 or tp % NG == 0
        self.norm_G = max(NG // tp, 1)
        self.group_ranks = max(tp // NG, 1)                         # ranks sharing one norm group (statistics are combined over them)
        self.full_group_width = I // NG
        h0, g0 = r * self.nh, (r * G) // tp

        def rows(t, width, start, count, base):                     # rows [base + start*width, base + (start+count)*width)
            return t[base + start * width: base + (start + count) * width]

        def shard_in(full, rank):                                   # [gate | x | B | C | dt] -> local slices, same order
            hh, gg = rank * self.nh, (rank * G) // tp
            return torch.cat([rows(full, hd, hh, self.nh, 0), rows(full, hd, hh, self.nh, I), rows(full, N, gg, self.G, 2 * I),
                              rows(full, N, gg, self.G, 2 * I + G * N), rows(full, 1, hh, self.nh, 2 * I + 2 * G * N)], 0).contiguous()

        def shard_conv(full, rank):                                 # [x | B | C]
            hh, gg = rank * self.nh, (rank * G) // tp
            return torch.cat([rows(full, hd, hh, self.nh, 0), rows(full, N, gg, self.G, I), rows(full, N, gg, self.G, I + G * N)], 0).contiguous()

        def mk(*shape, shard=None):
            p = nn.Parameter(torch.zeros(*shape, dtype=dt, device=device), requires_grad=False)
            p.partition_dim, p.tp_group = 0, g
            if shard is not None:
                p.shard_fn = shard
            return p
        heads = lambda full, rank: full[rank * self.nh: (rank + 1) * self.nh].contiguous()                       # noqa: E731
        chans = lambda full, rank: full[rank * self.I: (rank + 1) * self.I].contiguous()                         # noqa: E731
        self.in_proj_weight = mk(self.I + self.conv_dim + self.nh, H, shard=shard_in)
        self.in_proj_bias = mk(self.I + self.conv_dim + self.nh, shard=shard_in) if getattr(config, "mamba_proj_bias", False) else None
        ob = bool(getattr(config, "projectors_bias", False)) if out_bias is None else bool(out_bias)
        self.out_proj = RowParallelLinear(I, H, bias=ob, input_is_parallel=True, dtype=dt, device=device)
        self.conv_weight = mk(self.conv_dim, self.K, shard=shard_conv)
        self.conv_bias = mk(self.conv_dim, shard=shard_conv) if getattr(config, "mamba_conv_bias", True) else None
        self.dt_bias, self.A_log, self.D = mk(self.nh, shard=heads), mk(self.nh, shard=heads), mk(self.nh, shard=heads)
        self.gated_norm = bool(getattr(config, "mamba_rms_norm", False)) if gated_norm is None else bool(gated_norm)
        self.norm_before_gate = bool(getattr(config, "mamba_norm_before_gate", True)) if norm_before_gate is None else bool(norm_before_gate)
        if self.gated_norm:
            self.norm_weight = mk(self.I, shard=chans)
        self.eps = config.rms_norm_eps
        lim = getattr(config, "time_step_limit", (0.0, float("inf")))
        self.dt_min, self.dt_max = float(lim[0]), float(lim[1])
        self.conv_state, self.ssm_state = f"m2_conv{i}", f"m2_ssm{i}"
        del h0, g0

    def state_specs(self):
        return {self.conv_state: (self.K - 1, self.conv_dim), self.ssm_state: ((self.nh, self.hd, self.N), torch.float32)}

    def _group_rms(self, y, B, T):
        """RMS-normalise every group of ``full_group_width`` channels (local share: ``I / G_local``)."""
        yg = y.view(B, T, self.norm_G, self.I // self.norm_G)
        ss = yg.pow(2).sum(-1, keepdim=True)
        if self.group_ranks > 1:                                    # my group is spread over ``group_ranks`` consecutive ranks
            from ...parallel import mappings
            allss = mappings.all_gather(ss.unsqueeze(0).contiguous(), 0, self.tp_group)                          # [tp, B, T, 1, 1]
            first = (self.tp_group.rank // self.group_ranks) * self.group_ranks
            ss = allss[first: first + self.group_ranks].sum(0)
        return (yg * torch.rsqrt(ss / self.full_group_width + self.eps)).reshape(B, T, self.I)

    def forward(self, xn, meta, kv_mgr):
        B, T, _ = xn.shape
        K, nh, hd, G, N = self.K, self.nh, self.hd, self.G, self.N
        gate, xBC, dt = ops.linear(xn, self.in_proj_weight, self.in_proj_bias).split([self.I, self.conv_dim, nh], -1)
        lines = kv_mgr.lines_for(meta.seq_ids)
        states = kv_mgr.states
        w = self.conv_weight.t().unsqueeze(0)
        if meta.is_prefill:
            if meta.has_prefix:
                raise NotImplementedError("Mamba-2 with a cached prefix")
            n = _last_valid(meta, B, T, xn.device)
            pad = F.pad(xBC, (0, 0, K - 1, 0))
            conv = sum(pad[:, j:j + T] * w[:, j:j + 1] for j in range(K))
            idx = (n.view(B, 1) + torch.arange(K - 1, device=xn.device).view(1, -1)).unsqueeze(-1).expand(B, K - 1, self.conv_dim)
            states.write(self.conv_state, lines, pad.gather(1, idx))
            valid = torch.arange(T, device=xn.device).view(1, T) < n.view(B, 1)
            h = torch.zeros(B, nh, hd, N, dtype=torch.float32, device=xn.device)
        else:
            if T != 1:
                raise NotImplementedError("Mamba-2 takes one new token per decode step")
            win = torch.cat([states.read(self.conv_state, lines).to(xBC.dtype), xBC], 1)
            conv = (win * w).sum(1, keepdim=True)
            states.write(self.conv_state, lines, win[:, 1:])
            valid = torch.ones(B, 1, dtype=torch.bool, device=xn.device)
            h = states.read(self.ssm_state, lines).float()
        if self.conv_bias is not None:
            conv = conv + self.conv_bias
        x, Bm, Cm = F.silu(conv).split([self.I, G * N, G * N], -1)
        dt = F.softplus(dt.float() + self.dt_bias.float()).clamp(self.dt_min, self.dt_max)                   # [B, T, nh]
        dt = torch.where(valid.unsqueeze(-1), dt, torch.zeros_like(dt))                                          # padding: state carried through
        A = -torch.exp(self.A_log.float())
        xh = x.float().view(B, T, nh, hd)
        rep = nh // G
        Bh = Bm.float().view(B, T, G, N).repeat_interleave(rep, 2)                                               # [B, T, nh, N]
        Ch = Cm.float().view(B, T, G, N).repeat_interleave(rep, 2)
        dA = torch.exp(dt * A)                                                                                    # [B, T, nh]
        ys = []
        for t in range(T):
            h = h * dA[:, t, :, None, None] + (dt[:, t, :, None] * xh[:, t])[..., None] * Bh[:, t, :, None, :]
            ys.append((h * Ch[:, t, :, None, :]).sum(-1))
        states.write(self.ssm_state, lines, h)
        y = (torch.stack(ys, 1) + xh * self.D.float().view(1, 1, nh, 1)).reshape(B, T, self.I)
        if self.gated_norm:
            if not self.norm_before_gate:
                y = y * F.silu(gate.float())
            y = self._group_rms(y, B, T) * self.norm_weight.float()
            if self.norm_before_gate:
                y = y * F.silu(gate.float())
        else:
            y = y * F.silu(gate.float())
        return self.out_proj(y.to(xn.dtype))


class FalconH1Layer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        b = bool(getattr(config, "attention_bias", False))
        self.mamba = Mamba2Mixer(config, i, device)
        self.self_attn = NeuronLlamaAttention(config, i, rotary, device=device, qkv_bias=b, o_bias=b)
        self.mlp = GatedMLP(config.hidden_size, config.intermediate_size, config.hidden_act, dt, bias=bool(getattr(config, "mlp_bias", False)),
                            device=device)
        self.input_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.pre_ff_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.layer_idx = i

    def state_specs(self):
        return self.mamba.state_specs()

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.input_layernorm
        m = self.mamba(n(h), meta, kv_mgr)
        h = self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h) + m
        n = self.pre_ff_layernorm
        return self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)


class NeuronFalconH1Model(_HybridModel):
    def make_layer(self, config, i, rotary, device):
        return FalconH1Layer(config, i, rotary, device)

    def init_model(self, config):
        super().init_model(config)
        self.embed_scale = float(getattr(config, "embedding_multiplier", 1.0))


class NeuronFalconH1ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronFalconH1Model

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        """Every muP multiplier scales the input or the output of a linear map: fold them into the weights once."""
        I = config.mamba_d_ssm if getattr(config, "mamba_d_ssm", None) is not None else int(config.mamba_expand * config.hidden_size)
        gn = config.mamba_n_groups * config.mamba_d_state
        z = [float(v) for v in getattr(config, "ssm_multipliers", [1.0] * 5)]
        mup = torch.cat([torch.full((I,), z[0]), torch.full((I,), z[1]), torch.full((gn,), z[2]), torch.full((gn,), z[3]),
                         torch.full((config.mamba_n_heads,), z[4])])
        g = lambda name, d=1.0: float(getattr(config, name, d))                                                  # noqa: E731
        gate_m, down_m = [float(v) for v in getattr(config, "mlp_multipliers", [1.0, 1.0])]
        sc = lambda t, f: (t.float() * f).to(t.dtype)                                                             # noqa: E731
        out = {}
        for k, v in sd.items():
            k = k.replace(".feed_forward.", ".mlp.").replace("final_layernorm.", "norm.")
            if k.endswith(".mamba.in_proj.weight"):
                k, v = k.replace(".in_proj.weight", ".in_proj_weight"), sc(v, mup.view(-1, 1) * g("ssm_in_multiplier"))
            elif k.endswith(".mamba.in_proj.bias"):
                k, v = k.replace(".in_proj.bias", ".in_proj_bias"), sc(v, mup)
            elif k.endswith(".mamba.out_proj.weight") or k.endswith(".mamba.out_proj.bias"):
                v = sc(v, g("ssm_out_multiplier"))
            elif k.endswith(".mamba.conv1d.weight"):
                k, v = k.replace(".conv1d.weight", ".conv_weight"), v.squeeze(1)
            elif k.endswith(".mamba.conv1d.bias"):
                k = k.replace(".conv1d.bias", ".conv_bias")
            elif k.endswith(".mamba.norm.weight"):
                k = k.replace(".mamba.norm.weight", ".mamba.norm_weight")
            elif ".self_attn.q_proj.weight" in k or ".self_attn.v_proj.weight" in k:
                v = sc(v, g("attention_in_multiplier"))
            elif ".self_attn.k_proj.weight" in k:
                v = sc(v, g("attention_in_multiplier") * g("key_multiplier"))
            elif ".self_attn.k_proj.bias" in k:
                v = sc(v, g("key_multiplier"))
            elif ".self_attn.o_proj." in k:
                v = sc(v, g("attention_out_multiplier"))
            elif ".mlp.gate_proj." in k:
                v = sc(v, gate_m)
            elif ".mlp.down_proj." in k:
                v = sc(v, down_m)
            elif k == "lm_head.weight":
                v = sc(v, g("lm_head_multiplier"))
            out[k] = v
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = sc(out["embed_tokens.weight"], g("lm_head_multiplier"))
        return fuse_qkv_and_gate_up(out, config.num_hidden_layers)

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        pass


# ---------------------------------------------------------------------------------------------------------------------- Bamba
class BambaLayer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        dt = config.neuron_config.torch_dtype
        kinds = getattr(config, "layers_block_type", None)
        self.is_attn = (kinds[i] == "attention") if kinds else (i in (getattr(config, "attn_layer_indices", None) or []))
        if self.is_attn:
            b = bool(getattr(config, "attention_bias", False))
            self.self_attn = NeuronLlamaAttention(config, i, rotary, device=device, qkv_bias=b, o_bias=b)
        else:
            pb = bool(getattr(config, "mamba_proj_bias", False))
            self.mamba = Mamba2Mixer(config, i, device, gated_norm=True, norm_before_gate=False, norm_groups=1, out_bias=pb)
        self.mlp = GatedMLP(config.hidden_size, config.intermediate_size, config.hidden_act, dt, bias=bool(getattr(config, "mlp_bias", False)),
                            device=device)
        self.input_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.pre_ff_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.layer_idx = i

    def state_specs(self):
        return {} if self.is_attn else self.mamba.state_specs()

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.input_layernorm
        if self.is_attn:
            h = self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)
        else:
            h = h + self.mamba(n(h), meta, kv_mgr)
        n = self.pre_ff_layernorm
        return self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)


class BambaInferenceConfig(LlamaInferenceConfig):
    def add_derived_config(self):
        if getattr(self, "mamba_d_ssm", None) is None:
            self.mamba_d_ssm = int(getattr(self, "mamba_expand", 2) * self.hidden_size)
        super().add_derived_config()


class NeuronBambaModel(_HybridModel):
    def make_rotary(self, config, device):
        from ...models.llama.modeling_llama import rope_scaling_of, rope_theta_of
        from ...modules.rope import RotaryEmbedding
        rp = getattr(config, "rope_parameters", None) or {}
        frac = (rp.get("partial_rotary_factor") if isinstance(rp, dict) else None) or getattr(config, "partial_rotary_factor", None) or 1.0
        return RotaryEmbedding(int(config.head_dim * float(frac)), max(config.max_position_embeddings, config.neuron_config.seq_len),
                               rope_theta_of(config), rope_scaling_of(config), device=device)

    def make_layer(self, config, i, rotary, device):
        return BambaLayer(config, i, rotary, device)


class NeuronBambaForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronBambaModel

    @classmethod
    def get_config_cls(cls):
        return BambaInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            k = k.replace(".feed_forward.", ".mlp.").replace("final_layernorm.", "norm.")
            if k.endswith(".mamba.in_proj.weight"):
                k = k.replace(".in_proj.weight", ".in_proj_weight")
            elif k.endswith(".mamba.in_proj.bias"):
                k = k.replace(".in_proj.bias", ".in_proj_bias")
            elif k.endswith(".mamba.conv1d.weight"):
                k, v = k.replace(".conv1d.weight", ".conv_weight"), v.squeeze(1)
            elif k.endswith(".mamba.conv1d.bias"):
                k = k.replace(".conv1d.bias", ".conv_bias")
            elif k.endswith(".mamba.norm.weight"):
                k = k.replace(".mamba.norm.weight", ".mamba.norm_weight")
            out[k] = v
        return fuse_qkv_and_gate_up(out, config.num_hidden_layers)


# ---------------------------------------------------------------------------------------------------------------------- Granite-4.0 hybrid
class GraniteHybridLayer(nn.Module):
    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        from ...modules.moe import ExpertMLPs, MoE, RouterTopK, SharedExperts
        dt, H = config.neuron_config.torch_dtype, config.hidden_size
        self.is_attn = config.layers_block_type[i] != "mamba"
        if self.is_attn:
            b = bool(getattr(config, "attention_bias", False))
            self.self_attn = NeuronLlamaAttention(config, i, rotary, device=device, qkv_bias=b, o_bias=b,
                                                  use_rope=getattr(config, "position_embedding_type", "rope") == "rope",
                                                  softmax_scale=float(getattr(config, "attention_multiplier", None) or config.head_dim ** -0.5))
        else:
            pb = bool(getattr(config, "mamba_proj_bias", False))
            self.mamba = Mamba2Mixer(config, i, device, gated_norm=True, norm_before_gate=False, norm_groups=1, out_bias=pb)
        n_exp = int(getattr(config, "num_local_experts", 0) or 0)
        self.mlp_is_moe = n_exp > 0
        if self.mlp_is_moe:
            self.mlp = MoE(RouterTopK(n_exp, config.num_experts_per_tok, H, dt, "softmax", False, True, False, device),
                           ExpertMLPs(n_exp, H, config.intermediate_size, config.hidden_act, dt, device=device),
                           SharedExperts(H, config.shared_intermediate_size, config.hidden_act, dt, device))
        else:
            self.mlp = GatedMLP(H, config.shared_intermediate_size, config.hidden_act, dt, device=device)
        self.input_layernorm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.post_attention_layernorm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.layer_idx = i

    def state_specs(self):
        return {} if self.is_attn else self.mamba.state_specs()

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.input_layernorm
        if self.is_attn:
            h = self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)
        else:
            h = h + self.mamba(n(h), meta, kv_mgr)
        n = self.post_attention_layernorm
        return self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)


class GraniteHybridInferenceConfig(BambaInferenceConfig):
    @classmethod
    def get_neuron_config_cls(cls):
        from ...config import MoENeuronConfig
        return MoENeuronConfig


class NeuronGraniteHybridModel(_HybridModel):
    def make_layer(self, config, i, rotary, device):
        return GraniteHybridLayer(config, i, rotary, device)

    def init_model(self, config):
        super().init_model(config)
        self.embed_scale = float(getattr(config, "embedding_multiplier", 1.0))


class NeuronGraniteHybridForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronGraniteHybridModel

    @classmethod
    def get_config_cls(cls):
        return GraniteHybridInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        """The residual multiplier scales the OUTPUT of every residual branch (mixer out-projection, shared MLP and expert down
        projections) and the logits scaling divides the head: folded into those weights, as for dense Granite."""
        rm, ls = float(getattr(config, "residual_multiplier", 1.0)), float(getattr(config, "logits_scaling", 1.0))
        moe = int(getattr(config, "num_local_experts", 0) or 0) > 0
        sc = lambda t, f: (t.float() * f).to(t.dtype)                                             # noqa: E731
        out = {}
        for k, v in sd.items():
            if k.endswith(".mamba.in_proj.weight"):
                k = k.replace(".in_proj.weight", ".in_proj_weight")
            elif k.endswith(".mamba.in_proj.bias"):
                k = k.replace(".in_proj.bias", ".in_proj_bias")
            elif k.endswith(".mamba.conv1d.weight"):
                k, v = k.replace(".conv1d.weight", ".conv_weight"), v.squeeze(1)
            elif k.endswith(".mamba.conv1d.bias"):
                k = k.replace(".conv1d.bias", ".conv_bias")
            elif k.endswith(".mamba.norm.weight"):
                k = k.replace(".mamba.norm.weight", ".mamba.norm_weight")
            elif ".mamba.out_proj." in k or ".self_attn.o_proj." in k:
                v = sc(v, rm)
            elif k.endswith(".shared_mlp.input_linear.weight"):
                k = k.replace(".shared_mlp.input_linear.weight", ".mlp.shared_experts.gate_up_proj.weight" if moe else ".mlp.gate_up_proj.weight")
            elif k.endswith(".shared_mlp.output_linear.weight"):
                k, v = k.replace(".shared_mlp.output_linear.weight", ".mlp.shared_experts.down_proj.weight" if moe else ".mlp.down_proj.weight"), sc(v, rm)
            elif k.endswith(".block_sparse_moe.input_linear.weight"):
                k = k.replace(".block_sparse_moe.input_linear.weight", ".mlp.expert_mlps.gate_up_proj")
            elif k.endswith(".block_sparse_moe.output_linear.weight"):
                k, v = k.replace(".block_sparse_moe.output_linear.weight", ".mlp.expert_mlps.down_proj"), sc(v, rm)
            elif k.endswith(".block_sparse_moe.router.layer.weight"):
                k, v = k.replace(".block_sparse_moe.router.layer.weight", ".mlp.router.linear_router.weight"), v.float()
            out[k] = v
        out = fuse_qkv_and_gate_up(out, config.num_hidden_layers, fuse_mlp=False)
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        out["lm_head.weight"] = sc(out["lm_head.weight"], 1.0 / ls)
        return out

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        pass


# ---------------------------------------------------------------------------------------------------------------------- Mamba-2
class Mamba2InferenceConfig(LlamaInferenceConfig):
    attribute_map = {"num_heads": "mamba_n_heads", "state_size": "mamba_d_state", "n_groups": "mamba_n_groups",
                     "conv_kernel": "mamba_d_conv", "expand": "mamba_expand", "layer_norm_epsilon": "rms_norm_eps",
                     "use_conv_bias": "mamba_conv_bias", "use_bias": "mamba_proj_bias"}

    def get_required_attributes(self):
        return ["hidden_size", "num_hidden_layers", "vocab_size", "mamba_n_heads", "mamba_d_state"]

    def add_derived_config(self):
        self.mamba_d_head = self.head_dim                                   # the checkpoint's head_dim is the SSM head width
        self.mamba_d_ssm = self.mamba_n_heads * self.mamba_d_head
        self.num_attention_heads = self.num_key_value_heads = 1            # no attention: placeholders for the KV-cache plumbing
        self.head_dim, self.intermediate_size = 8, self.hidden_size
        self.max_position_embeddings = getattr(self, "max_position_embeddings", None) or self.neuron_config.seq_len
        if not hasattr(self, "hidden_act") or self.hidden_act is None:
            self.hidden_act = "silu"
        super().add_derived_config()


class Mamba2Layer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, device=None):
        super().__init__()
        pb = bool(getattr(config, "mamba_proj_bias", False))
        self.mixer = Mamba2Mixer(config, i, device, gated_norm=bool(getattr(config, "rms_norm", True)), norm_before_gate=False, norm_groups=1,
                                 out_bias=pb)
        self.norm = RMSNorm(config.hidden_size, config.rms_norm_eps, config.neuron_config.torch_dtype, device=device)
        self.layer_idx = i

    def state_specs(self):
        return self.mixer.state_specs()

    def forward(self, h, meta, kv_mgr, lora=None):
        return h + self.mixer(self.norm(h), meta, kv_mgr)


class NeuronMamba2Model(_HybridModel):
    def make_layer(self, config, i, rotary, device):
        return Mamba2Layer(config, i, device)


class NeuronMamba2ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronMamba2Model
    _STATE_DICT_MODEL_PREFIX = "backbone."

    @classmethod
    def get_config_cls(cls):
        return Mamba2InferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            k = k.replace("embeddings.", "embed_tokens.").replace("norm_f.", "norm.")
            if k.endswith(".mixer.in_proj.weight"):
                k = k.replace(".in_proj.weight", ".in_proj_weight")
            elif k.endswith(".mixer.in_proj.bias"):
                k = k.replace(".in_proj.bias", ".in_proj_bias")
            elif k.endswith(".mixer.conv1d.weight"):
                k, v = k.replace(".conv1d.weight", ".conv_weight"), v.squeeze(1)
            elif k.endswith(".mixer.conv1d.bias"):
                k = k.replace(".conv1d.bias", ".conv_bias")
            elif k.endswith(".mixer.norm.weight"):
                k = k.replace(".mixer.norm.weight", ".mixer.norm_weight")
            out[k] = v
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


# ---------------------------------------------------------------------------------------------------------------------- Nemotron-H
class _SharedPlainMLP(nn.Module):
    """Non-gated always-on expert (``down(act(up(x)))``); output left un-reduced like :class:`SharedExperts`."""

    def __init__(self, hidden_size, intermediate_size, act, dtype, device=None):
        super().__init__()
        from ...modules.mlp import _PLAIN_ACT
        self.act = _PLAIN_ACT[act]
        self.up_proj = ColumnParallelLinear(hidden_size, intermediate_size, bias=False, gather_output=False, dtype=dtype, device=device)
        self.down_proj = RowParallelLinear(intermediate_size, hidden_size, bias=False, dtype=dtype, device=device, reduce_output=False)

    def forward(self, x2, reduce: bool = False):
        return self.down_proj(ops.activation(self.up_proj(x2), self.act))


class NemotronHInferenceConfig(LlamaInferenceConfig):
    """NVIDIA Nemotron-H / Nemotron-3-Nano: one mixer per block — Mamba-2, attention WITHOUT positions, squared-ReLU MLP or a
    DeepSeek-style sigmoid-routed MoE of NON-gated squared-ReLU experts plus one always-on shared expert."""
    attribute_map = {"mamba_num_heads": "mamba_n_heads", "mamba_head_dim": "mamba_d_head", "n_groups": "mamba_n_groups",
                     "ssm_state_size": "mamba_d_state", "conv_kernel": "mamba_d_conv", "layer_norm_epsilon": "rms_norm_eps",
                     "use_conv_bias": "mamba_conv_bias", "use_bias": "mamba_proj_bias"}
    _PATTERN = {"M": "mamba", "*": "attention", "-": "mlp", "E": "moe"}

    def get_required_attributes(self):
        return ["hidden_size", "vocab_size", "num_attention_heads", "mamba_n_heads", "mamba_d_head", "mamba_d_state"]

    def add_derived_config(self):
        if not getattr(self, "layers_block_type", None):                    # original checkpoints: "M-M-M*-..." pattern string
            self.layers_block_type = [self._PATTERN[c] for c in self.hybrid_override_pattern]
        self.num_hidden_layers = len(self.layers_block_type)
        self.mamba_d_ssm = self.mamba_n_heads * self.mamba_d_head
        self.time_step_limit = (float(getattr(self, "time_step_min", 0.0) or 0.0), float("inf"))    # HF floors dt at time_step_min, no ceiling
        self.hidden_act = getattr(self, "mlp_hidden_act", "relu2")
        if getattr(self, "moe_latent_size", None) is not None:
            raise NotImplementedError("Nemotron-H latent-projected experts (moe_latent_size)")
        if getattr(self, "mlp_bias", False):
            raise NotImplementedError("Nemotron-H with MLP biases")
        super().add_derived_config()

    @classmethod
    def get_neuron_config_cls(cls):
        from ...config import MoENeuronConfig
        return MoENeuronConfig


class NemotronHLayer(nn.Module):
    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        from ...models.deepseek.modeling_deepseek import DeepseekRouter
        from ...modules.mlp import PlainMLP
        from ...modules.moe import ExpertMLPs, MoE
        dt, H, act = config.neuron_config.torch_dtype, config.hidden_size, config.hidden_act
        self.kind = config.layers_block_type[i]
        self.mlp_is_moe = self.kind == "moe"
        if self.kind == "attention":
            b = bool(getattr(config, "attention_bias", False))
            self.self_attn = NeuronLlamaAttention(config, i, rotary, device=device, qkv_bias=b, o_bias=b, use_rope=False)
        elif self.kind == "mamba":
            self.mamba = Mamba2Mixer(config, i, device, gated_norm=True, norm_before_gate=False, norm_groups=config.mamba_n_groups,
                                     out_bias=bool(getattr(config, "mamba_proj_bias", False)))
        elif self.kind == "moe":
            experts = ExpertMLPs(config.n_routed_experts, H, config.moe_intermediate_size, act, dt, device=device, gated=False)
            self.mlp = MoE(DeepseekRouter(config, device), experts,
                           _SharedPlainMLP(H, config.moe_shared_expert_intermediate_size, act, dt, device))
        elif self.kind == "mlp":
            self.mlp = PlainMLP(H, config.intermediate_size, act, dt, bias=False, device=device)
        else:
            raise ValueError(f"unknown Nemotron-H block type {self.kind!r}")
        self.norm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.layer_idx = i

    def state_specs(self):
        return self.mamba.state_specs() if self.kind == "mamba" else {}

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.norm
        if self.kind == "attention":
            return self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)
        if self.kind == "mamba":
            return h + self.mamba(n(h), meta, kv_mgr)
        if self.kind == "moe":
            return self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)
        return self.mlp(n(h), residual=h)


class NeuronNemotronHModel(_HybridModel):
    def make_layer(self, config, i, rotary, device):
        return NemotronHLayer(config, i, rotary, device)


class NeuronNemotronHForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronNemotronHModel

    @classmethod
    def get_config_cls(cls):
        return NemotronHInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        kinds, out = config.layers_block_type, {}
        ren = {"mamba": [(".mixer.in_proj.weight", ".mamba.in_proj_weight"), (".mixer.in_proj.bias", ".mamba.in_proj_bias"),
                         (".mixer.conv1d.weight", ".mamba.conv_weight"), (".mixer.conv1d.bias", ".mamba.conv_bias"),
                         (".mixer.norm.weight", ".mamba.norm_weight"), (".mixer.", ".mamba.")],
               "attention": [(".mixer.", ".self_attn.")],
               "mlp": [(".mixer.up_proj.", ".mlp.fc1."), (".mixer.down_proj.", ".mlp.fc2.")],
               "moe": [(".mixer.gate.weight", ".mlp.router.linear_router.weight"),
                       (".mixer.gate.e_score_correction_bias", ".mlp.router.e_score_correction_bias"),
                       (".mixer.experts.up_proj", ".mlp.expert_mlps.gate_up_proj"), (".mixer.experts.down_proj", ".mlp.expert_mlps.down_proj"),
                       (".mixer.shared_experts.", ".mlp.shared_experts.")]}
        sd = {(k[len("backbone."):] if k.startswith("backbone.") else k): v for k, v in sd.items()}   # NVIDIA's on-disk prefix
        for i, kind in enumerate(kinds):                                    # on disk the experts are separate Linear modules
            for proj in ("up_proj", "down_proj") if kind == "moe" else ():
                ws = [sd.pop(f"layers.{i}.mixer.experts.{e}.{proj}.weight", None) for e in range(config.n_routed_experts)]
                if ws[0] is not None:
                    sd[f"layers.{i}.mixer.experts.{proj}"] = torch.stack(ws)
        for k, v in sd.items():
            k = k.replace("embeddings.", "embed_tokens.").replace("embedding.", "embed_tokens.").replace("norm_f.", "norm.")
            if k.startswith("layers.") and ".mixer." in k:
                kind = kinds[int(k.split(".")[1])]
                for a, b in ren[kind]:
                    if a in k:
                        k = k.replace(a, b)
                        break
                if k.endswith(".mamba.conv_weight"):
                    v = v.squeeze(1)
                if ".mlp.router." in k:
                    v = v.float()
            out[k] = v
        out = fuse_qkv_and_gate_up(out, config.num_hidden_layers, fuse_mlp=False)
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out

    @staticmethod
    def update_state_dict_for_tied_weights(sd):
        pass


# ---------------------------------------------------------------------------------------------------------------------- Mamba-1
class Mamba1Mixer(nn.Module):
    """Selective SSM of Mamba-1 (per-CHANNEL ``dt`` and state ``[I, N]``): ``in_proj -> (x, gate)``, causal depthwise conv + SiLU on
    x, ``x_proj -> (dt_low_rank, B, C)``, ``dt = softplus(dt_proj(.))``, ``h_t = exp(dt A) h_{t-1} + dt B_t x_t``,
    ``y = C_t h_t + D x``, gated by ``silu(gate)``.  Tensor parallel over the inner channels: ``x_proj`` is row-parallel (its small
    ``R + 2N`` output is all-reduced), everything else is channel-local.  ``bcdt_rms``: Falcon-Mamba normalises B, C and the low-rank
    dt (weight-free RMS)."""

    def __init__(self, config, i, device=None, bcdt_rms=False):
        super().__init__()
        dt, H = config.neuron_config.torch_dtype, config.hidden_size
        I, self.N, self.K, self.R = config.mamba_d_inner, config.mamba_d_state, config.mamba_d_conv, config.mamba_dt_rank
        pb = bool(getattr(config, "mamba_proj_bias", False))
        self.in_proj = ColumnParallelLinear(H, 2 * I, bias=pb, gather_output=False, dtype=dt, device=device, stride=2)
        self.x_proj = RowParallelLinear(I, self.R + 2 * self.N, bias=False, input_is_parallel=True, dtype=dt, device=device)
        self.dt_proj = ColumnParallelLinear(self.R, I, bias=True, gather_output=False, dtype=dt, device=device)
        self.out_proj = RowParallelLinear(I, H, bias=pb, input_is_parallel=True, dtype=dt, device=device)
        g = self.in_proj.tensor_parallel_group
        self.I = I // g.size

        def mk(*shape):
            p = nn.Parameter(torch.zeros(*shape, dtype=dt, device=device), requires_grad=False)
            p.partition_dim, p.tp_group = 0, g
            return p
        self.conv_weight = mk(self.I, self.K)
        self.conv_bias = mk(self.I) if getattr(config, "mamba_conv_bias", True) else None
        self.A_log, self.D = mk(self.I, self.N), mk(self.I)
        self.bcdt_eps = float(getattr(config, "mixer_rms_eps", 1e-6)) if bcdt_rms else None
        if bcdt_rms == "weighted":                                 # Jamba: learned RMSNorms on the low-rank dt, B and C (replicated)
            self.bcdt_eps = None
            self.dt_layernorm = RMSNorm(self.R, config.rms_norm_eps, dt, device=device)
            self.b_layernorm = RMSNorm(self.N, config.rms_norm_eps, dt, device=device)
            self.c_layernorm = RMSNorm(self.N, config.rms_norm_eps, dt, device=device)
        self.weighted_norms = bcdt_rms == "weighted"
        self.conv_state, self.ssm_state = f"m1_conv{i}", f"m1_ssm{i}"

    def state_specs(self):
        return {self.conv_state: (self.K - 1, self.I), self.ssm_state: ((self.I, self.N), torch.float32)}

    def forward(self, xn, meta, kv_mgr):
        B, T, _ = xn.shape
        K, N, R = self.K, self.N, self.R
        x, gate = self.in_proj(xn).chunk(2, -1)
        lines, states = kv_mgr.lines_for(meta.seq_ids), kv_mgr.states
        w = self.conv_weight.t().unsqueeze(0)
        if meta.is_prefill:
            if meta.has_prefix:
                raise NotImplementedError("Mamba with a cached prefix")
            n = _last_valid(meta, B, T, xn.device)
            pad = F.pad(x, (0, 0, K - 1, 0))
            conv = sum(pad[:, j:j + T] * w[:, j:j + 1] for j in range(K))
            idx = (n.view(B, 1) + torch.arange(K - 1, device=xn.device).view(1, -1)).unsqueeze(-1).expand(B, K - 1, self.I)
            states.write(self.conv_state, lines, pad.gather(1, idx))
            valid = torch.arange(T, device=xn.device).view(1, T) < n.view(B, 1)
            h = torch.zeros(B, self.I, N, dtype=torch.float32, device=xn.device)
        else:
            if T != 1:
                raise NotImplementedError("Mamba takes one new token per decode step")
            win = torch.cat([states.read(self.conv_state, lines).to(x.dtype), x], 1)
            conv = (win * w).sum(1, keepdim=True)
            states.write(self.conv_state, lines, win[:, 1:])
            valid = torch.ones(B, 1, dtype=torch.bool, device=xn.device)
            h = states.read(self.ssm_state, lines).float()
        if self.conv_bias is not None:
            conv = conv + self.conv_bias
        u = F.silu(conv)
        dtr, Bm, Cm = self.x_proj(u).split([R, N, N], -1)
        if self.bcdt_eps is not None:
            rms = lambda t: (t.float() * torch.rsqrt(t.float().pow(2).mean(-1, keepdim=True) + self.bcdt_eps)).to(t.dtype)     # noqa: E731
            dtr, Bm, Cm = rms(dtr), rms(Bm), rms(Cm)
        elif self.weighted_norms:
            dtr, Bm, Cm = self.dt_layernorm(dtr), self.b_layernorm(Bm), self.c_layernorm(Cm)
        dt = F.softplus(self.dt_proj(dtr).float())                                                   # [B, T, I]
        dt = torch.where(valid.unsqueeze(-1), dt, torch.zeros_like(dt))                              # padding: state carried through
        A = -torch.exp(self.A_log.float())                                                            # [I, N]
        uf, Bf, Cf = u.float(), Bm.float(), Cm.float()
        ys = []
        for t in range(T):
            h = h * torch.exp(dt[:, t, :, None] * A) + (dt[:, t] * uf[:, t])[..., None] * Bf[:, t, None, :]
            ys.append((h * Cf[:, t, None, :]).sum(-1))
        states.write(self.ssm_state, lines, h)
        y = (torch.stack(ys, 1) + uf * self.D.float()) * F.silu(gate.float())
        return self.out_proj(y.to(xn.dtype))


class Mamba1InferenceConfig(LlamaInferenceConfig):
    attribute_map = {"state_size": "mamba_d_state", "conv_kernel": "mamba_d_conv", "time_step_rank": "mamba_dt_rank",
                     "layer_norm_epsilon": "rms_norm_eps", "use_conv_bias": "mamba_conv_bias", "use_bias": "mamba_proj_bias"}

    def get_required_attributes(self):
        return ["hidden_size", "num_hidden_layers", "vocab_size", "mamba_d_state", "mamba_d_conv"]

    def add_derived_config(self):
        import math
        self.mamba_d_inner = int(getattr(self, "intermediate_size", None) or getattr(self, "expand", 2) * self.hidden_size)
        if isinstance(self.mamba_dt_rank, str) or self.mamba_dt_rank is None:                        # "auto"
            self.mamba_dt_rank = math.ceil(self.hidden_size / 16)
        self.num_attention_heads = self.num_key_value_heads = 1                                     # no attention: KV-cache plumbing placeholders
        self.head_dim, self.intermediate_size = 8, self.mamba_d_inner
        self.max_position_embeddings = getattr(self, "max_position_embeddings", None) or self.neuron_config.seq_len
        self.hidden_act = getattr(self, "hidden_act", None) or "silu"
        super().add_derived_config()


class Mamba1Layer(nn.Module):
    mlp_is_moe = False

    def __init__(self, config, i, device=None, bcdt_rms=False):
        super().__init__()
        self.mixer = Mamba1Mixer(config, i, device, bcdt_rms=bcdt_rms)
        self.norm = RMSNorm(config.hidden_size, config.rms_norm_eps, config.neuron_config.torch_dtype, device=device)
        self.layer_idx = i

    def state_specs(self):
        return self.mixer.state_specs()

    def forward(self, h, meta, kv_mgr, lora=None):
        return h + self.mixer(self.norm(h), meta, kv_mgr)


class NeuronMambaModel(_HybridModel):
    bcdt_rms = False

    def make_layer(self, config, i, rotary, device):
        return Mamba1Layer(config, i, device, bcdt_rms=self.bcdt_rms)


class NeuronFalconMambaModel(NeuronMambaModel):
    bcdt_rms = True


class NeuronMambaForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronMambaModel
    _STATE_DICT_MODEL_PREFIX = "backbone."

    @classmethod
    def get_config_cls(cls):
        return Mamba1InferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            k = k.replace("embeddings.", "embed_tokens.").replace("norm_f.", "norm.")
            if k.endswith(".mixer.conv1d.weight"):
                k, v = k.replace(".conv1d.weight", ".conv_weight"), v.squeeze(1)
            elif k.endswith(".mixer.conv1d.bias"):
                k = k.replace(".conv1d.bias", ".conv_bias")
            out[k] = v
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


class NeuronFalconMambaForCausalLM(NeuronMambaForCausalLM):
    _model_cls = NeuronFalconMambaModel


# ---------------------------------------------------------------------------------------------------------------------- Jamba
class JambaInferenceConfig(LlamaInferenceConfig):
    def add_derived_config(self):
        import math
        L = self.num_hidden_layers
        self.mamba_d_inner = int(getattr(self, "mamba_expand", 2) * self.hidden_size)
        if isinstance(getattr(self, "mamba_dt_rank", "auto"), str):
            self.mamba_dt_rank = math.ceil(self.hidden_size / 16)
        if not getattr(self, "layers_block_type", None):
            p, o = getattr(self, "attn_layer_period", 8), getattr(self, "attn_layer_offset", 4)
            self.layers_block_type = ["attention" if i % p == o else "mamba" for i in range(L)]
        if not getattr(self, "layers_num_experts", None):
            p, o = getattr(self, "expert_layer_period", 2), getattr(self, "expert_layer_offset", 1)
            self.layers_num_experts = [self.num_experts if i % p == o else 1 for i in range(L)]
        super().add_derived_config()

    @classmethod
    def get_neuron_config_cls(cls):
        from ...config import MoENeuronConfig
        return MoENeuronConfig


class JambaLayer(nn.Module):
    """AI21 Jamba: Mamba-1 (with learned dt / B / C norms) or position-free GQA attention, then a SwiGLU that is a softmax top-k MoE
    (weights NOT renormalised) on every ``expert_layer_period``-th layer."""

    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        from ...modules.moe import ExpertMLPs, MoE, RouterTopK
        dt, H = config.neuron_config.torch_dtype, config.hidden_size
        self.is_attn = config.layers_block_type[i] == "attention"
        if self.is_attn:
            self.self_attn = NeuronLlamaAttention(config, i, rotary, device=device, use_rope=False)
        else:
            self.mamba = Mamba1Mixer(config, i, device, bcdt_rms="weighted")
        E = int(config.layers_num_experts[i])
        self.mlp_is_moe = E > 1
        if self.mlp_is_moe:
            self.mlp = MoE(RouterTopK(E, config.num_experts_per_tok, H, dt, "softmax", False, False, False, device),
                           ExpertMLPs(E, H, config.intermediate_size, config.hidden_act, dt, device=device))
        else:
            self.mlp = GatedMLP(H, config.intermediate_size, config.hidden_act, dt, device=device)
        self.input_layernorm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.pre_ff_layernorm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.layer_idx = i

    def state_specs(self):
        return {} if self.is_attn else self.mamba.state_specs()

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.input_layernorm
        if self.is_attn:
            h = self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)
        else:
            h = h + self.mamba(n(h), meta, kv_mgr)
        n = self.pre_ff_layernorm
        return self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)


class NeuronJambaModel(_HybridModel):
    def make_layer(self, config, i, rotary, device):
        return JambaLayer(config, i, rotary, device)


class NeuronJambaForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronJambaModel

    @classmethod
    def get_config_cls(cls):
        return JambaInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        from ...models.state_dict_utils import convert_moe_experts
        out = {}
        for k, v in sd.items():
            k = k.replace("final_layernorm.", "norm.")
            if k.endswith(".mamba.conv1d.weight"):
                k, v = k.replace(".conv1d.weight", ".conv_weight"), v.squeeze(1)
            elif k.endswith(".mamba.conv1d.bias"):
                k = k.replace(".conv1d.bias", ".conv_bias")
            out[k] = v
        moe_layers = [i for i, e in enumerate(config.layers_num_experts) if int(e) > 1]
        out = convert_moe_experts(out, config.num_hidden_layers, config.num_experts, moe_prefixes=("feed_forward",), gate_names=("router",),
                                  w_names=("gate_proj", "up_proj", "down_proj"), layers=moe_layers)
        out = {k.replace(".feed_forward.", ".mlp."): v for k, v in out.items()}
        out = fuse_qkv_and_gate_up(out, config.num_hidden_layers)
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


HYBRID_MODEL_TYPES = {"jamba": NeuronJambaForCausalLM, "mamba": NeuronMambaForCausalLM, "falcon_mamba": NeuronFalconMambaForCausalLM, "nemotron_h": NeuronNemotronHForCausalLM, "lfm2_moe": NeuronLfm2MoeForCausalLM, "mamba2": NeuronMamba2ForCausalLM, "granitemoehybrid": NeuronGraniteHybridForCausalLM, "bamba": NeuronBambaForCausalLM, "falcon_h1": NeuronFalconH1ForCausalLM, "lfm2": NeuronLfm2ForCausalLM, "recurrent_gemma": NeuronRecurrentGemmaForCausalLM}
