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
