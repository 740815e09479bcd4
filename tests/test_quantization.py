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
