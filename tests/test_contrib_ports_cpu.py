"""Contrib ports of architectures that exist on the hub only as remote code (MiniCPM4, InternLM3, Orion) or as the text backbone of
a multimodal checkpoint (Janus, Ovis2.5).  Oracle: the Hugging Face class that implements the same maths, with the checkpoint's
config.json / weight names rewritten to the port's layout."""
import json
import os

import pytest
import torch

from neuronx_distributed_inference_b200.config import load_pretrained_config
from neuronx_distributed_inference_b200.contrib.models.backbone_ports import PORT_MODEL_TYPES
from neuronx_distributed_inference_b200.utils.accuracy import generate_expected_logits, teacher_forced_logits
from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint

BASE = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2, vocab_size=160,
            max_position_embeddings=256)


def _rewrite_config(ckpt, **changes):
    """config.json only: the weights file stays untouched (the oracle memory-maps it)."""
    p = os.path.join(ckpt, "config.json")
    cfg = json.load(open(p))
    drop = changes.pop("_drop", ())
    cfg.update(changes)
    for k in drop:
        cfg.pop(k, None)
    json.dump(cfg, open(p, "w"))


def _check(name, hf, ckpt):
    cls = PORT_MODEL_TYPES[name]
    nc = cls.get_neuron_config_cls()(batch_size=2, seq_len=48, max_context_length=24, torch_dtype="float32", on_cpu=True, output_logits=True)
    app = cls(ckpt, cls.get_config_cls()(nc, load_config=load_pretrained_config(ckpt)))
    app.load(None, skip_warmup=True)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(1, 160, (2, 14), generator=g)
    mask = torch.ones_like(ids)
    mask[1, 10:] = 0
    exp, toks = generate_expected_logits(hf, ids, mask, 8)
    got = teacher_forced_logits(app, ids, mask, toks)
    err = ((got - exp).norm() / exp.norm()).item()
    assert err < 3e-4, f"{name}: relative logit error {err}"


def _hf(cfg, path):
    from transformers import AutoModelForCausalLM
    ckpt = save_random_hf_checkpoint(cfg, path, seed=3)
    return AutoModelForCausalLM.from_pretrained(ckpt, dtype=torch.float32).eval(), ckpt


def test_minicpm_is_granite_with_mup_names(tmp_path):
    import transformers as T
    L, H = 3, 64
    hf, ckpt = _hf(T.GraniteConfig(**BASE, embedding_multiplier=12.0, residual_multiplier=1.4 / L ** 0.5, logits_scaling=H / 32,
                                   attention_multiplier=16 ** -0.5, tie_word_embeddings=False), str(tmp_path / "m"))
    _rewrite_config(ckpt, model_type="minicpm", architectures=["MiniCPMForCausalLM"], scale_emb=12.0, scale_depth=1.4, dim_model_base=32,
                    _drop=("embedding_multiplier", "residual_multiplier", "logits_scaling", "attention_multiplier"))
    _check("minicpm", hf, ckpt)


def test_internlm3_bias_switches(tmp_path):
    import transformers as T
    hf, ckpt = _hf(T.LlamaConfig(**BASE, attention_bias=True, mlp_bias=True), str(tmp_path / "i"))
    _rewrite_config(ckpt, model_type="internlm3", architectures=["InternLM3ForCausalLM"], qkv_bias=True, bias=True,
                    _drop=("attention_bias", "mlp_bias"))
    _check("internlm3", hf, ckpt)


def test_orion_layernorm_llama(tmp_path):
    import transformers as T
    hf, ckpt = _hf(T.StableLmConfig(**BASE, partial_rotary_factor=1.0, use_qkv_bias=False, layer_norm_eps=1e-5), str(tmp_path / "o"))
    _rewrite_config(ckpt, model_type="orion", architectures=["OrionForCausalLM"], rms_norm_eps=1e-5,
                    _drop=("layer_norm_eps", "partial_rotary_factor", "use_qkv_bias", "rope_parameters"))
    _check("orion", hf, ckpt)


@pytest.mark.parametrize("name,prefix,nest", [("janus", "model.language_model.", "text_config"), ("ovis2_5", "llm.model.", "llm_config")])
def test_text_backbone_of_multimodal_checkpoint(name, prefix, nest, tmp_path):
    """Causal-LM weights moved under the multimodal prefix, hyper-parameters nested under ``text_config`` / ``llm_config``, plus a
    vision tensor the port has to ignore.  Written to a NEW directory: the oracle's weights are memory-mapped from the original."""
    import transformers as T
    from safetensors.torch import load_file, save_file
    cfg = T.LlamaConfig(**BASE) if name == "janus" else T.Qwen3Config(**BASE, head_dim=16)
    hf, ckpt = _hf(cfg, str(tmp_path / name))
    outer = prefix[: -len("model.")] if prefix.endswith(".model.") else prefix          # where lm_head lives
    new = {}
    for k, v in load_file(os.path.join(ckpt, "model.safetensors")).items():
        new[(prefix + k[len("model."):]) if k.startswith("model.") else (outer + k)] = v.clone()
    new["vision_model.dummy.weight"] = torch.zeros(2, 2)
    dst = str(tmp_path / (name + "_mm"))
    os.makedirs(dst)
    save_file(new, os.path.join(dst, "model.safetensors"))
    inner = json.load(open(os.path.join(ckpt, "config.json")))
    json.dump({"model_type": "composite", nest: inner, "vision_config": {"hidden_size": 8}}, open(os.path.join(dst, "config.json"), "w"))
    _check(name, hf, dst)


def test_qwen2_5_omni_thinker_text_backbone(tmp_path):
    """The thinker's text decoder out of a full Omni checkpoint layout (``thinker.model.*``, ``thinker_config.text_config``);
    the oracle is the Hugging Face thinker itself on a text-only prompt."""
    import transformers as T
    from safetensors.torch import save_file
    from transformers.models.qwen2_5_omni import modeling_qwen2_5_omni as M
    torch.manual_seed(0)
    tc = T.Qwen2_5OmniThinkerConfig(
        text_config=dict(**BASE, rope_parameters=dict(rope_type="default", mrope_section=[2, 3, 3], rope_theta=10000.0)),
        audio_config=dict(d_model=16, encoder_layers=1, encoder_attention_heads=2, encoder_ffn_dim=32, num_mel_bins=8, output_dim=64),
        vision_config=dict(depth=1, hidden_size=16, intermediate_size=32, num_heads=2, out_hidden_size=64, patch_size=4, spatial_merge_size=2,
                           temporal_patch_size=2))
    hf = M.Qwen2_5OmniThinkerForConditionalGeneration(tc).eval()
    dst = str(tmp_path / "omni")
    os.makedirs(dst)
    save_file({"thinker." + k: v.clone().contiguous() for k, v in hf.state_dict().items()}, os.path.join(dst, "model.safetensors"))
    json.dump({"model_type": "qwen2_5_omni_test", "thinker_config": tc.to_dict()}, open(os.path.join(dst, "config.json"), "w"))
    _check("qwen2_5_omni", hf, dst)


def test_nemotron_h_pattern_string_and_dense_mlp_block(tmp_path):
    """Original Nemotron-H checkpoints describe the stack with ``hybrid_override_pattern`` ("M" Mamba-2, "*" attention, "-" squared-ReLU
    MLP, "E" MoE) and transformers 5.5 cannot instantiate the "-" blocks, so the pattern parsing and the dense block are checked here
    directly (the Mamba-2 / attention / MoE blocks are checked against HF in test_contrib_cpu)."""
    import transformers as T
    from neuronx_distributed_inference_b200.contrib.models.hybrid_family import NemotronHInferenceConfig, NemotronHLayer
    hf = T.NemotronHConfig(hidden_size=64, intermediate_size=96, layers_block_type=["mamba", "attention"], num_attention_heads=4,
                           num_key_value_heads=2, head_dim=16, vocab_size=160, mamba_num_heads=8, mamba_head_dim=16, n_groups=2,
                           ssm_state_size=8, conv_kernel=4, n_routed_experts=4, moe_intermediate_size=32,
                           moe_shared_expert_intermediate_size=48, max_position_embeddings=256)
    hf.save_pretrained(str(tmp_path))
    path = os.path.join(str(tmp_path), "config.json")
    raw = json.load(open(path))
    raw.pop("layers_block_type", None)
    raw["hybrid_override_pattern"] = "M-*E"
    json.dump(raw, open(path, "w"))
    nc = NemotronHInferenceConfig.get_neuron_config_cls()(batch_size=1, seq_len=32, torch_dtype="float32", on_cpu=True)

    def load(cfg):                                    # AutoConfig would re-derive layers_block_type: feed the raw dictionary
        for k, v in raw.items():
            setattr(cfg, k, v)
    cfg = NemotronHInferenceConfig(nc, load_config=load)
    assert cfg.layers_block_type == ["mamba", "mlp", "attention", "moe"] and cfg.num_hidden_layers == 4
    assert cfg.mamba_d_ssm == 128 and cfg.mamba_n_groups == 2 and cfg.hidden_act == "relu2"
    layer = NemotronHLayer(cfg, 1, None)
    torch.manual_seed(0)
    for p in layer.parameters():
        p.data.normal_(0, 0.2)
    h = torch.randn(1, 5, 64)
    n = layer.norm
    x = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + n.variance_epsilon) * n.weight
    ref = h + torch.relu(x @ layer.mlp.fc1.weight.T).square() @ layer.mlp.fc2.weight.T
    assert torch.allclose(layer(h, None, None), ref, atol=1e-5)
