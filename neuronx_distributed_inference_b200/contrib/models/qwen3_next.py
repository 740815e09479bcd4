"""Qwen3-Next: three Gated-DeltaNet (linear attention) layers per gated softmax-attention layer, each followed by a Qwen2-MoE style
sparse feed-forward (top-k routed SwiGLU experts + a sigmoid-gated shared expert); every RMSNorm is zero-centred (``1 + w``).

Gated DeltaNet, per value head (state ``S`` [d_k, d_v], fp32, one per cache line in the :class:`RecurrentStateCache`):
``S <- exp(g_t) S``; ``S <- S + k_t (beta_t (v_t - S^T k_t))^T``; ``o_t = S^T q_t / sqrt(d_k)`` with l2-normalised q / k,
``beta = sigmoid(b)``, ``g = -exp(A_log) * softplus(a + dt_bias)``; q / k / v come out of one projection followed by a causal depthwise
convolution + SiLU; the output is RMS-normalised per head, gated by ``silu(z)`` and projected back.
Tensor parallel over the KEY heads (their value heads, convolution channels and ``out_proj`` columns follow).
The softmax layers: per-head (1 + w) q / k RMSNorm, rotary on the first quarter of the head, sigmoid output gate computed by the second
half of ``q_proj``.  Checked against Hugging Face (tests/test_contrib_cpu.py)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...models.llama.modeling_llama import NeuronLlamaForCausalLM, NeuronLlamaMLP
from ...models.state_dict_utils import convert_moe_experts, fuse_qkv_and_gate_up
from ...modules.norm import RMSNorm
from ...parallel.layers import RowParallelLinear
from .hybrid_family import _HybridModel, _last_valid
from .moe_family import _AfmoeAttention, _is_moe_layer, _MoeConfig


class GatedDeltaNet(nn.Module):
    def __init__(self, config, i, device=None):
        super().__init__()
        from ...parallel.state import get_tensor_model_parallel_group
        dt, H = config.neuron_config.torch_dtype, config.hidden_size
        self.tp_group = g = get_tensor_model_parallel_group()
        tp = g.size
        nk, nv, dk, dv = config.linear_num_key_heads, config.linear_num_value_heads, config.linear_key_head_dim, config.linear_value_head_dim
        assert nk % tp == 0 and nv % nk == 0, "Gated DeltaNet key heads vs tp degree"
        self.nk, self.r, self.dk, self.dv, self.K = nk // tp, nv // nk, dk, dv, config.linear_conv_kernel_dim
        self.nv = self.nk * self.r
        self.key_dim, self.value_dim = self.nk * dk, self.nv * dv
        self.conv_dim = 2 * self.key_dim + self.value_dim
        KD, VD = nk * dk, nv * dv                                     # full widths (for slicing the checkpoint)

        def mk(*shape, shard=None):
            p = nn.Parameter(torch.zeros(*shape, dtype=dt, device=device), requires_grad=False)
            p.partition_dim, p.tp_group = 0, g                        # default: plain chunk of dim 0 (whole key-head groups)
            if shard is not None:
                p.shard_fn = shard
            return p

        def shard_conv(full, rank):                                   # channels are [q (KD) | k (KD) | v (VD)]
            q, k, v = full[:KD], full[KD:2 * KD], full[2 * KD:]
            return torch.cat([q.chunk(tp, 0)[rank], k.chunk(tp, 0)[rank], v.chunk(tp, 0)[rank]], 0).contiguous()
        self.in_proj_qkvz = mk(self.nk * (2 * dk + 2 * self.r * dv), H)
        self.in_proj_ba = mk(self.nk * 2 * self.r, H)
        self.conv_weight = mk(self.conv_dim, self.K, shard=shard_conv)
        self.dt_bias, self.A_log = mk(self.nv), mk(self.nv)
        self.norm_weight = nn.Parameter(torch.ones(dv, dtype=dt, device=device), requires_grad=False)
        self.out_proj = RowParallelLinear(VD, H, bias=False, input_is_parallel=True, dtype=dt, device=device)
        self.eps = config.rms_norm_eps
        self.conv_state, self.rec_state = f"gdn_conv{i}", f"gdn_state{i}"
        del VD

    def state_specs(self):
        return {self.conv_state: (self.K - 1, self.conv_dim), self.rec_state: ((self.nv, self.dk, self.dv), torch.float32)}

    def forward(self, xn, meta, kv_mgr):
        B, T, _ = xn.shape
        nk, r, dk, dv, K = self.nk, self.r, self.dk, self.dv, self.K
        qkvz = ops.linear(xn, self.in_proj_qkvz).view(B, T, nk, 2 * dk + 2 * r * dv)
        ba = ops.linear(xn, self.in_proj_ba).view(B, T, nk, 2 * r)
        q, k, v, z = qkvz.split([dk, dk, r * dv, r * dv], -1)
        b, a = ba.split([r, r], -1)
        mixed = torch.cat([q.reshape(B, T, -1), k.reshape(B, T, -1), v.reshape(B, T, -1)], -1)              # [B, T, conv_dim]
        lines, states = kv_mgr.lines_for(meta.seq_ids), kv_mgr.states
        w = self.conv_weight.t().unsqueeze(0)
        if meta.is_prefill:
            if meta.has_prefix:
                raise NotImplementedError("Gated DeltaNet with a cached prefix")
            n = _last_valid(meta, B, T, xn.device)
            pad = F.pad(mixed, (0, 0, K - 1, 0))
            conv = sum(pad[:, j:j + T] * w[:, j:j + 1] for j in range(K))
            idx = (n.view(B, 1) + torch.arange(K - 1, device=xn.device).view(1, -1)).unsqueeze(-1).expand(B, K - 1, self.conv_dim)
            states.write(self.conv_state, lines, pad.gather(1, idx))
            valid = torch.arange(T, device=xn.device).view(1, T) < n.view(B, 1)
            S = torch.zeros(B, self.nv, dk, dv, dtype=torch.float32, device=xn.device)
        else:
            if T != 1:
                raise NotImplementedError("Gated DeltaNet takes one new token per decode step")
            win = torch.cat([states.read(self.conv_state, lines).to(mixed.dtype), mixed], 1)
            conv = (win * w).sum(1, keepdim=True)
            states.write(self.conv_state, lines, win[:, 1:])
            valid = torch.ones(B, 1, dtype=torch.bool, device=xn.device)
            S = states.read(self.rec_state, lines).float()
        q, k, v = F.silu(conv).split([self.key_dim, self.key_dim, self.value_dim], -1)
        l2 = lambda t: t * torch.rsqrt(t.pow(2).sum(-1, keepdim=True) + 1e-6)                                 # noqa: E731
        qh = l2(q.float().view(B, T, nk, dk)).repeat_interleave(r, 2) * dk ** -0.5                           # [B, T, nv, dk]
        kh = l2(k.float().view(B, T, nk, dk)).repeat_interleave(r, 2)
        vh = v.float().view(B, T, self.nv, dv)
        beta = torch.sigmoid(b.float().reshape(B, T, self.nv))
        g = -torch.exp(self.A_log.float()) * F.softplus(a.float().reshape(B, T, self.nv) + self.dt_bias.float())
        # padded steps leave the state untouched: no decay, no write
        decay = torch.where(valid.unsqueeze(-1), torch.exp(g), torch.ones_like(g))
        beta = torch.where(valid.unsqueeze(-1), beta, torch.zeros_like(beta))
        outs = []
        for t in range(T):
            S = S * decay[:, t, :, None, None]
            kt = kh[:, t]
            delta = (vh[:, t] - (S * kt.unsqueeze(-1)).sum(-2)) * beta[:, t].unsqueeze(-1)
            S = S + kt.unsqueeze(-1) * delta.unsqueeze(-2)
            outs.append((S * qh[:, t].unsqueeze(-1)).sum(-2))
        states.write(self.rec_state, lines, S)
        o = torch.stack(outs, 1)                                                                               # [B, T, nv, dv]
        o = o * torch.rsqrt(o.pow(2).mean(-1, keepdim=True) + self.eps) * self.norm_weight.float()
        o = o * F.silu(z.float().reshape(B, T, self.nv, dv))
        return self.out_proj(o.reshape(B, T, self.value_dim).to(xn.dtype))


class _Qwen3NextAttention(_AfmoeAttention):
    """(1 + w) q / k norm is folded at load; rotary on every layer; the sigmoid output gate is the AFMoE mechanism."""

    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        from ...models.llama.modeling_llama import NeuronLlamaAttention
        NeuronLlamaAttention.__init__(self, config, layer_idx, rotary_emb, device=device, qk_norm="rms_pre_rope", qk_norm_eps=config.rms_norm_eps,
                                      qkv_bias=bool(getattr(config, "attention_bias", False)), o_bias=bool(getattr(config, "attention_bias", False)),
                                      **over)
        dt, D = config.neuron_config.torch_dtype, self.head_dim
        plan = self.qkv_proj.plan
        self.gate_weight = nn.Parameter(torch.zeros(self.n_q * D, config.hidden_size, dtype=dt, device=device), requires_grad=False)
        from ...modules.gqa import _gather_heads
        self.gate_weight.shard_fn = lambda full, rank: _gather_heads(full, plan.q_idx[rank], D, 0)
        self.gate_weight.partition_dim, self.gate_weight.tp_group = 0, self.tp_group
        self._gate = None


class Qwen3NextLayer(nn.Module):
    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        from ...modules.moe import SharedExperts, initialize_moe_module
        dt, H = config.neuron_config.torch_dtype, config.hidden_size
        self.is_attn = config.layer_types[i] == "full_attention"
        if self.is_attn:
            self.self_attn = _Qwen3NextAttention(config, i, rotary, device=device)
        else:
            self.linear_attn = GatedDeltaNet(config, i, device)
        self.mlp_is_moe = _is_moe_layer(config, i)
        if self.mlp_is_moe:
            self.mlp = initialize_moe_module(config, device=device, intermediate_size=config.moe_intermediate_size,
                                             normalize=bool(getattr(config, "norm_topk_prob", True)))
            self.mlp.shared_experts = SharedExperts(H, config.shared_expert_intermediate_size, config.hidden_act, dt, device)
            self.mlp.shared_expert_gate = nn.Linear(H, 1, bias=False, dtype=dt, device=device)
            self.mlp.shared_expert_gate.weight.requires_grad_(False)
        else:
            self.mlp = NeuronLlamaMLP(config, device=device)
        self.input_layernorm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.post_attention_layernorm = RMSNorm(H, config.rms_norm_eps, dt, device=device)
        self.layer_idx = i

    def state_specs(self):
        return {} if self.is_attn else self.linear_attn.state_specs()

    def forward(self, h, meta, kv_mgr, lora=None):
        n = self.input_layernorm
        if self.is_attn:
            h = self.self_attn(h, meta, kv_mgr, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)
        else:
            h = h + self.linear_attn(n(h), meta, kv_mgr)
        n = self.post_attention_layernorm
        return self.mlp(h, norm_weight=n.weight, norm_eps=n.variance_epsilon, residual=h)


class Qwen3NextInferenceConfig(_MoeConfig):
    def add_derived_config(self):
        if not getattr(self, "layer_types", None):
            k = getattr(self, "full_attention_interval", 4)
            self.layer_types = ["full_attention" if (i + 1) % k == 0 else "linear_attention" for i in range(self.num_hidden_layers)]
        rp = getattr(self, "rope_parameters", None) or {}
        if rp.get("partial_rotary_factor") is not None:
            self.partial_rotary_factor = float(rp["partial_rotary_factor"])
        super().add_derived_config()


class NeuronQwen3NextModel(_HybridModel):
    def make_layer(self, config, i, rotary, device):
        return Qwen3NextLayer(config, i, rotary, device)


class NeuronQwen3NextForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronQwen3NextModel

    @classmethod
    def get_config_cls(cls):
        return Qwen3NextInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        nq, D = config.num_attention_heads, config.head_dim
        out = {}
        for k, v in sd.items():
            if k.startswith("mtp."):                                                      # multi-token-prediction head: not used for decoding
                continue
            if k.endswith(".self_attn.q_proj.weight"):                                   # rows are [head, (query, gate), D]
                w = v.view(nq, 2, D, v.shape[-1])
                out[k.replace(".q_proj.weight", ".gate_weight")] = w[:, 1].reshape(nq * D, -1).contiguous()
                v = w[:, 0].reshape(nq * D, -1).contiguous()
            elif k.endswith((".input_layernorm.weight", ".post_attention_layernorm.weight", ".q_norm.weight", ".k_norm.weight")) or k == "norm.weight":
                v = (v.float() + 1.0).to(v.dtype)                                         # zero-centred RMSNorm
            elif k.endswith(".linear_attn.conv1d.weight"):
                k, v = k.replace(".conv1d.weight", ".conv_weight"), v.squeeze(1)
            elif k.endswith(".linear_attn.norm.weight"):
                k = k.replace(".norm.weight", ".norm_weight")
            elif k.endswith((".linear_attn.in_proj_qkvz.weight", ".linear_attn.in_proj_ba.weight")):
                k = k[:-len(".weight")]
            k = k.replace(".self_attn.q_norm.", ".self_attn.q_layernorm.").replace(".self_attn.k_norm.", ".self_attn.k_layernorm.")
            out[k] = v
        moe_layers = [i for i in range(config.num_hidden_layers) if _is_moe_layer(config, i)]
        out = fuse_qkv_and_gate_up(out, config.num_hidden_layers, fuse_mlp=True)
        out = convert_moe_experts(out, config.num_hidden_layers, config.num_experts, moe_prefixes=("mlp",), gate_names=("gate",),
                                  w_names=("gate_proj", "up_proj", "down_proj"), layers=moe_layers)
        out = {k.replace(".mlp.shared_expert.", ".mlp.shared_experts."): v for k, v in out.items()}
        for i in moe_layers:
            gk, uk = f"layers.{i}.mlp.shared_experts.gate_proj.weight", f"layers.{i}.mlp.shared_experts.up_proj.weight"
            if gk in out:
                out[f"layers.{i}.mlp.shared_experts.gate_up_proj.weight"] = torch.cat([out.pop(gk), out.pop(uk)], 0)
        if "lm_head.weight" not in out:
            out["lm_head.weight"] = out["embed_tokens.weight"].clone()
        return out


# ---- Qwen3.5 (text decoders): the same block; the DeltaNet projections are stored as four flat matrices ------------------------------------
class NeuronQwen3_5ForCausalLM(NeuronQwen3NextForCausalLM):
    """``qwen3_5_text`` / ``qwen3_5_moe_text`` (and the language model inside the ``qwen3_5`` / ``qwen3_5_moe`` checkpoints): in_proj_qkv
    ``[q | k | v]``, in_proj_z, in_proj_b, in_proj_a are regrouped per key head into the fused layout of :class:`GatedDeltaNet`."""
    _STATE_DICT_MODEL_PREFIX = "model.language_model."

    @classmethod
    def get_config_cls(cls):
        class Qwen3_5InferenceConfig(Qwen3NextInferenceConfig):
            def __init__(self, *a, **kw):
                self.intermediate_size = None                                             # all-MoE checkpoints do not state a dense width
                super().__init__(*a, **kw)

            def add_derived_config(self):
                if not hasattr(self, "num_experts") or self.num_experts is None:
                    self.num_experts, self.num_experts_per_tok = 0, 0                     # dense variant
                if not getattr(self, "intermediate_size", None):
                    self.intermediate_size = self.moe_intermediate_size
                super().add_derived_config()
        return Qwen3_5InferenceConfig

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        nk, nv, dk, dv = config.linear_num_key_heads, config.linear_num_value_heads, config.linear_key_head_dim, config.linear_value_head_dim
        r, KD = nv // nk, config.linear_num_key_heads * config.linear_key_head_dim
        sd = {(k[len("model."):] if k.startswith("model.") and not k.startswith("model.language_model.") else k): v for k, v in sd.items()
              if not k.startswith(("model.visual.", "visual.", "mtp."))}
        for i, kind in enumerate(config.layer_types):
            b = f"layers.{i}.linear_attn."
            if kind == "full_attention" or b + "in_proj_qkv.weight" not in sd:
                continue
            qkv, z = sd.pop(b + "in_proj_qkv.weight"), sd.pop(b + "in_proj_z.weight")
            H = qkv.shape[-1]
            q, k, v = qkv[:KD].view(nk, dk, H), qkv[KD:2 * KD].view(nk, dk, H), qkv[2 * KD:].view(nk, r * dv, H)
            sd[b + "in_proj_qkvz.weight"] = torch.cat([q, k, v, z.view(nk, r * dv, H)], 1).reshape(-1, H).contiguous()
            bb, aa = sd.pop(b + "in_proj_b.weight").view(nk, r, H), sd.pop(b + "in_proj_a.weight").view(nk, r, H)
            sd[b + "in_proj_ba.weight"] = torch.cat([bb, aa], 1).reshape(-1, H).contiguous()
        return NeuronQwen3NextForCausalLM.convert_hf_to_neuron_state_dict(sd, config)
