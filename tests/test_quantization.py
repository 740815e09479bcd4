"""Quantisation flow of README example 4: quantise the HF weights per channel, save, reload through
``quantized_checkpoints_path`` into converted modules, and check the quantised model tracks the fp32 one."""
import os

import pytest
import torch

from neuronx_distributed_inference_b200.config import NeuronConfig, load_pretrained_config
from neuronx_distributed_inference_b200.models.llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaForCausalLM
from neuronx_distributed_inference_b200.ops import reference as ref
from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint


def test_quantize_primitives_roundtrip():
    w = torch.randn(64, 96)
    for dt, tol in ((torch.int8, 0.02), (torch.float8_e4m3fn, 0.08)):
        q, s = ref.quantize_per_channel(w, dt)
        assert q.dtype == dt and s.shape == (64,)
        back = ref._dequant_weight(q, s)
        assert (back - w).abs().max() / w.abs().max() < tol
    q, s = ref.quantize_per_tensor(w, torch.int8)
    assert s.shape == (1,)
    q, s = ref.quantize_blockwise(torch.randn(256, 256), (128, 128), torch.float8_e4m3fn)
    assert s.shape == (2, 2)


@pytest.mark.parametrize("qtype,qdtype", [("per_channel_symmetric", "int8"), ("per_channel_symmetric", "f8e4m3"),
                                          ("per_tensor_symmetric", "int8")])
def test_llama_quantized_checkpoint_flow(tmp_path, qtype, qdtype):
    from transformers import LlamaConfig
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=160, max_position_embeddings=128)
    ckpt = save_random_hf_checkpoint(cfg, str(tmp_path / "ck"), seed=0)
    qpath = str(tmp_path / "q")

    def build(quantized):
        nc = NeuronConfig(batch_size=1, seq_len=32, max_context_length=16, torch_dtype="float32", on_cpu=True, output_logits=True,
                          quantized=quantized, quantized_checkpoints_path=qpath if quantized else None,
                          quantization_type=qtype, quantization_dtype=qdtype, modules_to_not_convert=["lm_head"])
        c = LlamaInferenceConfig(nc, load_config=load_pretrained_config(ckpt))
        return NeuronLlamaForCausalLM(ckpt, c), c
    qapp, qcfg = build(True)
    NeuronLlamaForCausalLM.save_quantized_state_dict(ckpt, qcfg)
    assert os.path.isdir(qpath)
    qapp.load(None, skip_warmup=True)
    attn = qapp.model.layers[0].self_attn
    assert attn.qkv_proj.weight.dtype in (torch.int8, torch.float8_e4m3fn) and attn.qkv_proj.scale.dtype == torch.float32
    assert qapp.model.lm_head.weight.dtype == torch.float32          # modules_to_not_convert honoured
    fapp, _ = build(False)
    fapp.load(None, skip_warmup=True)
    ids = torch.randint(1, 160, (1, 9))
    lq = qapp(ids).logits[:, -1].float()
    lf = fapp(ids).logits[:, -1].float()
    rel = ((lq - lf).norm() / lf.norm()).item()
    assert rel < (0.03 if qdtype == "int8" else 0.12), rel


def test_mxfp4_roundtrip_and_gpt_oss_conversion():
    import torch
    from neuronx_distributed_inference_b200.models.gpt_oss.mx_layout_transform import (dequantize_mxfp4, dequantize_mxfp4_state_dict,
                                                                                        pack_fp4_x4_uint16, quantize_mxfp4)
    torch.manual_seed(0)
    w = torch.randn(3, 8, 64)
    blocks, scales = quantize_mxfp4(w)
    assert blocks.shape == (3, 8, 2, 16) and scales.shape == (3, 8, 2)
    back = dequantize_mxfp4(blocks, scales, torch.float32)
    # e2m1 has 2 significant bits: relative error of a block is bounded by the grid spacing at its largest element
    assert (back - w).abs().max() <= 0.26 * w.abs().max()
    again = dequantize_mxfp4(*quantize_mxfp4(back), torch.float32)
    assert torch.equal(again, back)                                  # values on the grid are reproduced exactly
    sd = dequantize_mxfp4_state_dict({"layers.0.mlp.experts.down_proj_blocks": blocks, "layers.0.mlp.experts.down_proj_scales": scales})
    assert sd["layers.0.mlp.experts.down_proj"].shape == (3, 64, 8)
    assert pack_fp4_x4_uint16(torch.tensor([[1, 2, 3, 15]])).item() == 1 | (2 << 4) | (3 << 8) | (15 << 12)


def test_llama4_fp8_checkpoint_expert_fusion():
    """reference models/llama4/conversion_script/preprocess_llama4_FP8.py: per-expert fp8 weights + per-channel scales -> fused bf16
    ``experts.gate_up_proj [E, H, 2I]`` / ``down_proj [E, I, H]`` (and the --keep-fp8 variant with fused scales)."""
    import torch
    from neuronx_distributed_inference_b200.models.llama4.conversion_script.preprocess_llama4_fp8 import dequantize_dense, fuse_layer_experts
    torch.manual_seed(0)
    E, H, I = 3, 16, 8
    f8 = torch.float8_e4m3fn

    def q(shape):
        w = torch.randn(*shape)
        s = w.abs().amax(1, keepdim=True) / 448.0
        return (w / s).to(f8), s.to(torch.bfloat16)
    sd, ref = {}, {}
    p = "language_model.model.layers.1."
    for e in range(E):
        for name, shape in (("gate_proj", (I, H)), ("up_proj", (I, H)), ("down_proj", (H, I))):
            w, s = q(shape)
            sd[f"{p}feed_forward.experts.{e}.{name}.weight"], sd[f"{p}feed_forward.experts.{e}.{name}.weight_scale"] = w, s
            ref[(e, name)] = w.float() * s.float()
    w, s = q((I, H))
    sd[p + "feed_forward.shared_expert.gate_proj.weight"], sd[p + "feed_forward.shared_expert.gate_proj.weight_scale"] = w, s
    keep = {k: v.clone() for k, v in sd.items()}
    assert fuse_layer_experts(sd, p, E) and not fuse_layer_experts(sd, "language_model.model.layers.0.", E)
    gu, dn = sd[p + "feed_forward.experts.gate_up_proj"], sd[p + "feed_forward.experts.down_proj"]
    assert gu.shape == (E, H, 2 * I) and dn.shape == (E, I, H) and gu.dtype == torch.bfloat16
    x = torch.randn(5, H)
    for e in range(E):
        g, u = (x @ gu[e].float()).chunk(2, -1)
        exp_g, exp_u = x @ ref[(e, "gate_proj")].t(), x @ ref[(e, "up_proj")].t()
        assert torch.allclose(g, exp_g, atol=0.1, rtol=2e-2) and torch.allclose(u, exp_u, atol=0.1, rtol=2e-2)
        assert torch.allclose(dn[e].float(), ref[(e, "down_proj")].t(), atol=2e-2, rtol=2e-2)
    assert not any(".experts.0." in k for k in sd) and dequantize_dense(sd) == 1
    assert sd[p + "feed_forward.shared_expert.gate_proj.weight"].dtype == torch.bfloat16
    # fp8 kept: fused weights stay e4m3fn, scales are fused alongside and reproduce the same values
    assert fuse_layer_experts(keep, p, E, keep_fp8=True)
    gu8, gs = keep[p + "feed_forward.experts.gate_up_proj"], keep[p + "feed_forward.experts.gate_up_proj.scale"]
    assert gu8.dtype == f8 and gu8.shape == (E, H, 2 * I) and gs.shape == (E, 1, 2 * I)
    assert torch.allclose(gu8.float() * gs, gu.float(), atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("qdtype", ["int8", "f8e4m3"])
def test_mixtral_expert_wise_quantized_checkpoint_flow(qdtype, tmp_path):
    """``expert_wise_per_channel_symmetric``: the MoE expert banks are stored 8-bit with one scale per expert and output
    channel, attention projections per channel; the quantised model tracks the fp32 one."""
    from transformers import MixtralConfig
    from neuronx_distributed_inference_b200.models.mixtral.modeling_mixtral import NeuronMixtralForCausalLM as A
    from neuronx_distributed_inference_b200.quantization.convert import quantize_experts
    w = torch.randn(3, 8, 16)
    q, s = quantize_experts(w, torch.int8)
    assert q.dtype == torch.int8 and s.shape == (3, 8) and ((q.float() * s.unsqueeze(-1) - w).abs().max() / w.abs().max()) < 0.02
    cfg = MixtralConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=160,
                        max_position_embeddings=128, num_local_experts=4, num_experts_per_tok=2)
    ckpt = save_random_hf_checkpoint(cfg, str(tmp_path / "ck"), seed=0)
    qpath = str(tmp_path / "q")

    def build(quantized):
        nc = A.get_neuron_config_cls()(batch_size=1, seq_len=32, max_context_length=16, torch_dtype="float32", on_cpu=True, output_logits=True,
                                       quantized=quantized, quantized_checkpoints_path=qpath if quantized else None,
                                       quantization_type="expert_wise_per_channel_symmetric", quantization_dtype=qdtype,
                                       modules_to_not_convert=["lm_head"])
        c = A.get_config_cls()(nc, load_config=load_pretrained_config(ckpt))
        return A(ckpt, c), c
    qapp, qcfg = build(True)
    A.save_quantized_state_dict(ckpt, qcfg)
    qapp.load(None, skip_warmup=True)
    ex = qapp.model.layers[0].mlp.expert_mlps
    want = torch.int8 if qdtype == "int8" else torch.float8_e4m3fn
    assert ex.gate_up_proj.dtype == want and ex.down_proj.dtype == want
    assert ex.gate_up_scale.shape == (4, 256) and ex.down_scale.shape == (4, 64) and ex.gate_up_scale.dtype == torch.float32
    assert not torch.all(ex.gate_up_scale == 1) and qapp.model.layers[0].self_attn.qkv_proj.weight.dtype == want
    fapp, _ = build(False)
    fapp.load(None, skip_warmup=True)
    ids = torch.randint(1, 160, (1, 9))
    lq, lf = qapp(ids).logits[:, -1].float(), fapp(ids).logits[:, -1].float()
    rel = ((lq - lf).norm() / lf.norm()).item()
    assert rel < (0.03 if qdtype == "int8" else 0.12), rel
