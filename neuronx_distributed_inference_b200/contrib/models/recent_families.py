"""Families added after the reference hub was cut, each a few lines over the shared blocks:

* **Ministral / CWM (Code World Model)** — the Llama block with ``layer_types`` choosing sliding-window or full attention per layer
  (combinable with the window-sized rolling KV cache, ``rolling_sliding_window_cache``).
* **OLMo-1** — sequential pre-norm block with PARAMETER-FREE LayerNorms (unit weights are injected at load), SwiGLU, optional
  ``clip_qkv`` clamp on the fused projection output.
Each is checked against Hugging Face in fp32 (tests/test_contrib_cpu.py)."""
from __future__ import annotations

import torch

from ...models.llama.modeling_llama import NeuronLlamaAttention, NeuronLlamaForCausalLM, NeuronLlamaModel
from ...models.state_dict_utils import fuse_qkv_and_gate_up
from .classic_family import NeuronClassicModel, _ClassicCausalLM


class _LayerTypeAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        lt = getattr(config, "layer_types", None)
        sliding = lt[layer_idx] == "sliding_attention" if lt else getattr(config, "sliding_window", None) is not None
        b = bool(getattr(config, "attention_bias", False))
        super().__init__(config, layer_idx, rotary_emb, device=device, qkv_bias=b, o_bias=b,
                         sliding_window=getattr(config, "sliding_window", None) if sliding else None, **over)


class NeuronMinistralModel(NeuronLlamaModel):
    attention_cls = _LayerTypeAttention
    graph_safe = False


class NeuronMinistralForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronMinistralModel


NeuronCwmForCausalLM = NeuronMinistralForCausalLM


class NeuronOlmoModel(NeuronClassicModel):
    SPEC = dict(NeuronClassicModel.SPEC, norm_bias=False)

    def layer_spec(self, config, i):
        return dict(parallel=False, norm_bias=False, mlp="gated", act=getattr(config, "hidden_act", "silu"), qkv_bias=False, o_bias=False,
                    mlp_bias=False, clip_qkv=getattr(config, "clip_qkv", None))


class NeuronOlmoForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronOlmoModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers)
        ones = torch.ones(config.hidden_size, dtype=next(iter(sd.values())).dtype)
        for i in range(config.num_hidden_layers):
            sd[f"layers.{i}.input_layernorm.weight"] = ones.clone()
            sd[f"layers.{i}.post_attention_layernorm.weight"] = ones.clone()
        sd["norm.weight"] = ones.clone()
        return sd


RECENT_MODEL_TYPES = {"ministral": NeuronMinistralForCausalLM, "cwm": NeuronCwmForCausalLM, "olmo": NeuronOlmoForCausalLM}


# ---- HunYuan (dense and MoE): per-head q/k RMSNorm applied AFTER the rotation ---------------------------------------------------------
class _HunYuanAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        b = bool(getattr(config, "attention_bias", False))
        qk = "rms_post_rope" if getattr(config, "use_qk_norm", True) else None
        super().__init__(config, layer_idx, rotary_emb, device=device, qkv_bias=b, o_bias=b, qk_norm=qk, qk_norm_eps=config.rms_norm_eps, **over)


def _hunyuan_names(sd):
    return {k.replace(".self_attn.query_layernorm.", ".self_attn.q_layernorm.").replace(".self_attn.key_layernorm.", ".self_attn.k_layernorm."): v
            for k, v in sd.items()}


class NeuronHunYuanDenseModel(NeuronLlamaModel):
    attention_cls = _HunYuanAttention
    graph_safe = False


class NeuronHunYuanDenseForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronHunYuanDenseModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        return _hunyuan_names(NeuronLlamaForCausalLM.convert_hf_to_neuron_state_dict(sd, config))


RECENT_MODEL_TYPES["hunyuan_v1_dense"] = NeuronHunYuanDenseForCausalLM


class _HunYuanMoeModel(NeuronLlamaModel):
    """Softmax top-k (renormalised) routed SwiGLU experts plus an always-on shared SwiGLU of the dense width."""
    attention_cls = _HunYuanAttention
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        from ...models.model_base import DecoderLayer
        from ...modules.moe import ExpertMLPs, MoE, RouterTopK, SharedExperts
        from ...modules.norm import RMSNorm
        dt, H = config.neuron_config.torch_dtype, config.hidden_size
        pick = lambda v: v if isinstance(v, int) else v[i]                                          # noqa: E731  (per-layer lists allowed)
        E, k = pick(config.num_experts), pick(config.moe_topk)
        moe = MoE(RouterTopK(E, k, H, dt, "softmax", False, True, False, device),
                  ExpertMLPs(E, H, config.intermediate_size, config.hidden_act, dt, device=device),
                  SharedExperts(H, config.intermediate_size, config.hidden_act, dt, device))
        return DecoderLayer(self.attention_cls(config, i, rotary, device=device), moe, RMSNorm(H, config.rms_norm_eps, dt, device=device),
                            RMSNorm(H, config.rms_norm_eps, dt, device=device), i, mlp_is_moe=True)


class _MoeCfg:
    @classmethod
    def get_config_cls(cls):
        from .moe_family import _MoeConfig
        return _MoeConfig


class NeuronHunYuanMoEForCausalLM(_MoeCfg, NeuronLlamaForCausalLM):
    _model_cls = _HunYuanMoeModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        from ...models.state_dict_utils import convert_moe_experts
        sd = {k.replace(".mlp.gate.wg.", ".mlp.gate.").replace(".mlp.shared_mlp.", ".mlp.shared_experts."): v for k, v in sd.items()}
        for i in range(config.num_hidden_layers):                                                  # shared expert: [gate; up] fused rows
            b = f"layers.{i}.mlp.shared_experts."
            if b + "gate_proj.weight" in sd:
                sd[b + "gate_up_proj.weight"] = torch.cat([sd.pop(b + "gate_proj.weight"), sd.pop(b + "up_proj.weight")], 0)
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
        n_exp = config.num_experts if isinstance(config.num_experts, int) else max(config.num_experts)
        sd = convert_moe_experts(sd, config.num_hidden_layers, n_exp, moe_prefixes=("mlp",), gate_names=("gate",),
                                 w_names=("gate_proj", "up_proj", "down_proj"))
        return _hunyuan_names(sd)


# ---- FlexOlmo: the OLMo-2 block (norms AFTER attention / feed-forward, full-width q/k RMSNorm) with an OLMoE feed-forward ---------------
class FlexOlmoLayer(torch.nn.Module):
    mlp_is_moe = True

    def __init__(self, config, i, rotary, device=None):
        super().__init__()
        from ...modules.moe import initialize_moe_module
        from ...modules.norm import RMSNorm
        from .llama_family import Olmo2Attention
        dt = config.neuron_config.torch_dtype
        self.self_attn = Olmo2Attention(config, i, rotary, device)
        self.mlp = initialize_moe_module(config, device=device, intermediate_size=config.intermediate_size,
                                         normalize=bool(getattr(config, "norm_topk_prob", False)))
        self.post_attention_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.post_feedforward_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, dt, device=device)
        self.layer_idx = i

    def forward(self, h, meta, kv_mgr, lora=None):
        h = h + self.post_attention_layernorm(self.self_attn(h, meta, kv_mgr))
        return h + self.post_feedforward_layernorm(self.mlp(h))


class NeuronFlexOlmoModel(NeuronLlamaModel):
    graph_safe = False

    def make_layer(self, config, i, rotary, device):
        return FlexOlmoLayer(config, i, rotary, device)


class NeuronFlexOlmoForCausalLM(_MoeCfg, NeuronLlamaForCausalLM):
    _model_cls = NeuronFlexOlmoModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        from .moe_family import NeuronOlmoeForCausalLM
        return NeuronOlmoeForCausalLM.convert_hf_to_neuron_state_dict(sd, config)


# ---- GraniteMoE with the always-on shared SwiGLU (Granite-3.x MoE "shared" variants) -------------------------------------------------
def _granite_shared_model():
    from ...modules.moe import SharedExperts
    from .moe_family import NeuronGraniteMoeForCausalLM, NeuronGraniteMoeModel

    class NeuronGraniteMoeSharedModel(NeuronGraniteMoeModel):
        def make_layer(self, config, i, rotary, device):
            layer = super().make_layer(config, i, rotary, device)
            if getattr(config, "shared_intermediate_size", 0):
                layer.mlp.shared_experts = SharedExperts(config.hidden_size, config.shared_intermediate_size, config.hidden_act,
                                                         config.neuron_config.torch_dtype, device)
            return layer

    class NeuronGraniteMoeSharedForCausalLM(NeuronGraniteMoeForCausalLM):
        _model_cls = NeuronGraniteMoeSharedModel

        @staticmethod
        def convert_hf_to_neuron_state_dict(sd, config):
            rm = float(getattr(config, "residual_multiplier", 1.0))
            sd = dict(sd)
            for i in range(config.num_hidden_layers):
                a, b = f"layers.{i}.shared_mlp.", f"layers.{i}.mlp.shared_experts."
                if a + "input_linear.weight" in sd:
                    sd[b + "gate_up_proj.weight"] = sd.pop(a + "input_linear.weight")
                    w = sd.pop(a + "output_linear.weight")
                    sd[b + "down_proj.weight"] = (w.float() * rm).to(w.dtype)                      # residual multiplier folded, as for the experts
            return NeuronGraniteMoeForCausalLM.convert_hf_to_neuron_state_dict(sd, config)
    return NeuronGraniteMoeSharedForCausalLM


NeuronGraniteMoeSharedForCausalLM = _granite_shared_model()
RECENT_MODEL_TYPES.update({"hunyuan_v1_moe": NeuronHunYuanMoEForCausalLM, "flex_olmo": NeuronFlexOlmoForCausalLM,
                           "granitemoeshared": NeuronGraniteMoeSharedForCausalLM})


# ---- MiniMax-M2: full-width q/k RMSNorm, partial rotary, sigmoid router whose bias only steers the selection ---------------------------
class _SigmoidBiasRouter(torch.nn.Module):
    def __init__(self, num_experts, top_k, hidden_size, device=None):
        super().__init__()
        self.E, self.top_k = num_experts, top_k
        self.linear_router = torch.nn.Linear(hidden_size, num_experts, bias=False, dtype=torch.float32, device=device)
        self.linear_router.weight.requires_grad_(False)
        self.register_buffer("e_score_correction_bias", torch.zeros(num_experts, dtype=torch.float32, device=device))

    def forward(self, x):
        logits = torch.nn.functional.linear(x.float(), self.linear_router.weight)
        s = logits.sigmoid()
        idx = (s + self.e_score_correction_bias).topk(self.top_k, -1)[1]
        w = s.gather(1, idx)
        return logits, w / w.sum(-1, keepdim=True), idx


def _minimax_m2():
    from ...models.model_base import DecoderLayer
    from ...modules.moe import ExpertMLPs, MoE
    from ...modules.norm import RMSNorm
    from .llama_family import Olmo2Attention
    from .moe_family import _MoeConfig

    class MiniMaxM2Config(_MoeConfig):
        def add_derived_config(self):
            rp, rd = getattr(self, "rope_parameters", None) or {}, getattr(self, "rotary_dim", None)
            if rp.get("partial_rotary_factor") is not None:
                self.partial_rotary_factor = float(rp["partial_rotary_factor"])
            elif rd:                                        # original config.json: rotary_dim of head_dim channels rotate
                self.partial_rotary_factor = rd / (getattr(self, "head_dim", None) or self.hidden_size // self.num_attention_heads)
            super().add_derived_config()

    class NeuronMiniMaxM2Model(NeuronLlamaModel):
        graph_safe = False

        def make_layer(self, config, i, rotary, device):
            dt, H = config.neuron_config.torch_dtype, config.hidden_size
            attn = Olmo2Attention(config, i, rotary, device)
            moe = MoE(_SigmoidBiasRouter(config.num_local_experts, config.num_experts_per_tok, H, device),
                      ExpertMLPs(config.num_local_experts, H, config.intermediate_size, config.hidden_act, dt, device=device))
            return DecoderLayer(attn, moe, RMSNorm(H, config.rms_norm_eps, dt, device=device), RMSNorm(H, config.rms_norm_eps, dt, device=device),
                                i, mlp_is_moe=True)

    class NeuronMiniMaxM2ForCausalLM(NeuronLlamaForCausalLM):
        _model_cls = NeuronMiniMaxM2Model

        @classmethod
        def get_config_cls(cls):
            return MiniMaxM2Config

        @staticmethod
        def convert_hf_to_neuron_state_dict(sd, config):
            from ...models.state_dict_utils import convert_moe_experts
            sd = {k.replace(".block_sparse_moe.e_score_correction_bias", ".mlp.router.e_score_correction_bias"): v for k, v in sd.items()}
            sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
            sd = convert_moe_experts(sd, config.num_hidden_layers, config.num_local_experts, moe_prefixes=("block_sparse_moe",),
                                     gate_names=("gate",), w_names=("w1", "w3", "w2"))
            return {k.replace(".self_attn.q_norm.weight", ".self_attn.q_norm").replace(".self_attn.k_norm.weight", ".self_attn.k_norm"): v
                    for k, v in sd.items()}
    return NeuronMiniMaxM2ForCausalLM


NeuronMiniMaxM2ForCausalLM = _minimax_m2()
RECENT_MODEL_TYPES["minimax_m2"] = NeuronMiniMaxM2ForCausalLM


# ---- Solar-Open: the GLM-4.5-MoE block with every layer sparse and no q/k norm (DeepSeek-V3 sigmoid / group-limited router) -----------
def _solar_open():
    from .moe_family import NeuronGlm4MoeForCausalLM, _MoeConfig

    class SolarOpenConfig(_MoeConfig):
        def __init__(self, *a, **kw):
            self.intermediate_size = None                           # no dense layers: placeholder, set to the expert width below
            super().__init__(*a, **kw)

        def add_derived_config(self):
            if not getattr(self, "intermediate_size", None):
                self.intermediate_size = self.moe_intermediate_size
            super().add_derived_config()

    class NeuronSolarOpenForCausalLM(NeuronGlm4MoeForCausalLM):
        @classmethod
        def get_config_cls(cls):
            return SolarOpenConfig
    return NeuronSolarOpenForCausalLM


NeuronSolarOpenForCausalLM = _solar_open()
RECENT_MODEL_TYPES["solar_open"] = NeuronSolarOpenForCausalLM


# ---- EXAONE-MoE: EXAONE-4 attention (per-head q/k RMSNorm; hybrid models rotate only the sliding layers) in a pre-norm block whose
#      feed-forward is dense or a DeepSeek-V3-style MoE per ``mlp_layer_types`` --------------------------------------------------------
def _exaone_moe():
    from ...models.llama.modeling_llama import NeuronLlamaMLP
    from ...models.model_base import DecoderLayer
    from ...modules.attention import AttentionBase
    from ...modules.norm import RMSNorm
    from .moe_family import NeuronGlm4MoeForCausalLM, _deepseek_moe, _MoeConfig

    class ExaoneMoeConfig(_MoeConfig):
        def add_derived_config(self):
            self.n_routed_experts = self.num_experts
            self.n_shared_experts = getattr(self, "num_shared_experts", 0)
            if not getattr(self, "mlp_layer_types", None):
                self.mlp_layer_types = ["sparse"] * self.num_hidden_layers
            super().add_derived_config()

    class NeuronExaoneMoeModel(NeuronLlamaModel):
        graph_safe = False

        def make_layer(self, config, i, rotary, device):
            dt, H = config.neuron_config.torch_dtype, config.hidden_size
            lt = getattr(config, "layer_types", None)
            hybrid, sliding = getattr(config, "sliding_window", None) is not None, bool(lt and lt[i] == "sliding_attention")
            attn = AttentionBase(config, hidden_size=H, num_attention_heads=config.num_attention_heads,
                                 num_key_value_heads=config.num_key_value_heads, head_dim=config.head_dim, rotary_emb=rotary,
                                 qk_norm="rms_pre_rope", qk_norm_eps=config.rms_norm_eps, use_rope=(not hybrid) or sliding,
                                 sliding_window=config.sliding_window if (hybrid and sliding) else None, layer_idx=i,
                                 rms_norm_eps=config.rms_norm_eps, device=device)
            moe = config.mlp_layer_types[i] == "sparse"
            mlp = _deepseek_moe(config, device) if moe else NeuronLlamaMLP(config, device=device)
            return DecoderLayer(attn, mlp, RMSNorm(H, config.rms_norm_eps, dt, device=device), RMSNorm(H, config.rms_norm_eps, dt, device=device),
                                i, mlp_is_moe=moe)

    class NeuronExaoneMoeForCausalLM(NeuronGlm4MoeForCausalLM):
        _model_cls = NeuronExaoneMoeModel

        @classmethod
        def get_config_cls(cls):
            return ExaoneMoeConfig

        @staticmethod
        def convert_hf_to_neuron_state_dict(sd, config):
            sd = {k.replace(".mlp.e_score_correction_bias", ".mlp.gate.e_score_correction_bias"): v for k, v in sd.items()}   # on-disk spelling
            return NeuronGlm4MoeForCausalLM.convert_hf_to_neuron_state_dict(sd, config)
    return NeuronExaoneMoeForCausalLM


NeuronExaoneMoeForCausalLM = _exaone_moe()
RECENT_MODEL_TYPES["exaone_moe"] = NeuronExaoneMoeForCausalLM


# ---- Jais-2: LayerNorm (with bias) pre-norm block, rotary GQA, two-matrix MLP (squared ReLU by default), optional biases ---------------
class NeuronJais2Model(NeuronClassicModel):
    def layer_spec(self, config, i):
        ab, mb = bool(getattr(config, "attention_bias", True)), bool(getattr(config, "mlp_bias", True))
        return dict(parallel=False, norm_bias=True, mlp="plain", act=getattr(config, "hidden_act", "relu2"), qkv_bias=ab, o_bias=ab, mlp_bias=mb)


class NeuronJais2ForCausalLM(_ClassicCausalLM):
    _model_cls = NeuronJais2Model

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = {k.replace(".mlp.up_proj.", ".mlp.fc1.").replace(".mlp.down_proj.", ".mlp.fc2."): v for k, v in sd.items()}
        return fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)


RECENT_MODEL_TYPES["jais2"] = NeuronJais2ForCausalLM


# ---- MLA decoders that reuse the DeepSeek-V3 implementation: GLM-4.7-Flash (``glm4_moe_lite``) and Youtu-LLM (dense MLA) ---------------
def _mla_ports():
    from ...models.deepseek.modeling_deepseek import DeepseekInferenceConfig, NeuronDeepseekForCausalLM

    class Glm4MoeLiteConfig(DeepseekInferenceConfig):
        def add_derived_config(self):
            kinds = list(getattr(self, "mlp_layer_types", None) or [])
            if kinds:                                             # dense layers first, sparse after (the only layout DeepSeek-style stacks use)
                n_dense = kinds.index("sparse") if "sparse" in kinds else len(kinds)
                if any(k != "sparse" for k in kinds[n_dense:]):
                    raise NotImplementedError("dense layers after the first sparse layer")
                self.first_k_dense_replace = n_dense
            super().add_derived_config()

    class YoutuConfig(DeepseekInferenceConfig):
        def __init__(self, *a, **kw):
            self.n_routed_experts, self.num_experts_per_tok = 0, 0          # dense model: no expert layers at all
            super().__init__(*a, **kw)

        def add_derived_config(self):
            self.first_k_dense_replace = self.num_hidden_layers
            super().add_derived_config()

    class NeuronGlm4MoeLiteForCausalLM(NeuronDeepseekForCausalLM):
        @classmethod
        def get_config_cls(cls):
            return Glm4MoeLiteConfig

    class NeuronYoutuForCausalLM(NeuronDeepseekForCausalLM):
        @classmethod
        def get_config_cls(cls):
            return YoutuConfig
    return NeuronGlm4MoeLiteForCausalLM, NeuronYoutuForCausalLM


NeuronGlm4MoeLiteForCausalLM, NeuronYoutuForCausalLM = _mla_ports()
RECENT_MODEL_TYPES.update({"glm4_moe_lite": NeuronGlm4MoeLiteForCausalLM, "youtu": NeuronYoutuForCausalLM})


# ---- Ministral-3: YaRN rotary + the Llama-4 position-dependent query temperature  q *= 1 + beta * log(1 + floor(pos / L0)) -------------
class _Ministral3Attention(_LayerTypeAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, **over)
        rp = getattr(config, "rope_parameters", None) or {}
        self.q_beta = float(rp.get("llama_4_scaling_beta") or 0.0)
        self.q_l0 = float(rp.get("original_max_position_embeddings") or config.max_position_embeddings)

    def _simple(self):                      # the fused rope + append kernels do not scale q: take the generic path
        return self.q_beta == 0.0 and super()._simple()

    def _split_norm_rope(self, qkv, B, T, cos, sin, meta=None):
        q, k, v = super()._split_norm_rope(qkv, B, T, cos, sin, meta)
        if self.q_beta:
            pos = meta.position_ids if (meta is not None and meta.position_ids is not None) else torch.arange(T, device=q.device).view(1, T)
            temp = 1.0 + self.q_beta * torch.log1p(torch.floor(pos.float() / self.q_l0))
            q = q * temp.view(pos.shape[0], T, 1, 1).to(q.dtype)
        return q, k, v


class NeuronMinistral3Model(NeuronLlamaModel):
    attention_cls = _Ministral3Attention
    graph_safe = False


class NeuronMinistral3ForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronMinistral3Model


RECENT_MODEL_TYPES["ministral3"] = NeuronMinistral3ForCausalLM


# ---- nanochat: weight-free RMSNorms everywhere (embeddings, blocks, q / k after the rotation, output), squared-ReLU MLP, logit soft-cap --
class _NanoChatAttention(NeuronLlamaAttention):
    def __init__(self, config, layer_idx, rotary_emb, device=None, **over):
        super().__init__(config, layer_idx, rotary_emb, device=device, qk_norm="rms_post_rope", qk_norm_eps=config.rms_norm_eps, **over)

    def _rope(self, meta):                  # nanochat rotates the pairs the other way round: (x1, x2) -> (x1 c + x2 s, x2 c - x1 s)
        cos, sin = super()._rope(meta)
        return cos, -sin


class _NanoChatMLP(torch.nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        from ...modules.mlp import PlainMLP
        self.inner = PlainMLP(config.hidden_size, config.intermediate_size, config.hidden_act, config.neuron_config.torch_dtype, bias=False,
                              device=device)

    def forward(self, x, norm_weight=None, norm_eps=1e-6, norm_offset=0.0, residual=None, lora=None, adapter_ids=None):
        from ... import ops
        return self.inner(ops.rmsnorm(x, norm_weight, norm_eps, norm_offset) if norm_weight is not None else x, residual)


class NeuronNanoChatModel(NeuronLlamaModel):
    attention_cls = _NanoChatAttention
    mlp_cls = _NanoChatMLP
    graph_safe = False

    def init_model(self, config):
        super().init_model(config)
        self.final_logit_softcap = getattr(config, "final_logit_softcapping", None)
        self._embed_eps = config.rms_norm_eps

    def embed(self, input_ids, inputs_embeds=None, vision_embeddings=None, vision_mask=None):
        h = super().embed(input_ids, inputs_embeds, vision_embeddings, vision_mask)
        return (h.float() * torch.rsqrt(h.float().pow(2).mean(-1, keepdim=True) + self._embed_eps)).to(h.dtype)


class NeuronNanoChatForCausalLM(NeuronLlamaForCausalLM):
    _model_cls = NeuronNanoChatModel

    @staticmethod
    def convert_hf_to_neuron_state_dict(sd, config):
        sd = fuse_qkv_and_gate_up(sd, config.num_hidden_layers, fuse_mlp=False)
        sd = {k.replace(".mlp.fc", ".mlp.inner.fc"): v for k, v in sd.items()}
        dt = next(iter(sd.values())).dtype
        for i in range(config.num_hidden_layers):                               # the checkpoint has no norm weights at all: unit weights
            for n in ("input_layernorm", "post_attention_layernorm"):
                sd[f"layers.{i}.{n}.weight"] = torch.ones(config.hidden_size, dtype=dt)
            for n in ("q_layernorm", "k_layernorm"):
                sd[f"layers.{i}.self_attn.{n}.weight"] = torch.ones(config.head_dim, dtype=dt)
        sd["norm.weight"] = torch.ones(config.hidden_size, dtype=dt)
        return sd


RECENT_MODEL_TYPES["nanochat"] = NeuronNanoChatForCausalLM
