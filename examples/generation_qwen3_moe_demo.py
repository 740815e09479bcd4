"""Qwen3-MoE generation (reference examples/generation_qwen3_moe_demo.py): tensor parallel attention, experts sharded TP x EP.
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/generation_qwen3_moe_demo.py --model-path /path/to/Qwen3-30B-A3B \\
        --tp-degree 8 --moe-ep-degree 2 --moe-tp-degree 4"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import argparse

import torch

from neuronx_distributed_inference_b200.config import MoENeuronConfig, OnDeviceSamplingConfig, load_pretrained_config
from neuronx_distributed_inference_b200.models.qwen3_moe.modeling_qwen3_moe import NeuronQwen3MoeForCausalLM
from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--compiled-model-path", default="/tmp/qwen3_moe_artifacts")
    ap.add_argument("--tp-degree", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--moe-ep-degree", type=int, default=1)
    ap.add_argument("--moe-tp-degree", type=int, default=None)
    ap.add_argument("--prompt", action="append", default=None)
    ap.add_argument("--max-new-tokens", type=int, default=64)
    a = ap.parse_args()
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(a.model_path, padding_side="right")
    tok.pad_token = tok.pad_token or tok.eos_token
    prompts = a.prompt or ["Give me a short introduction to large language models.", "The capital of France is"]
    nc = MoENeuronConfig(tp_degree=a.tp_degree, moe_ep_degree=a.moe_ep_degree, moe_tp_degree=a.moe_tp_degree, batch_size=len(prompts),
                         max_context_length=256, seq_len=512, torch_dtype="bfloat16", on_device_sampling_config=OnDeviceSamplingConfig(top_k=1),
                         fused_qkv=True)
    cls = NeuronQwen3MoeForCausalLM
    model = cls(a.model_path, cls.get_config_cls()(nc, load_config=load_pretrained_config(a.model_path)))
    model.compile(a.compiled_model_path)
    model.load(a.compiled_model_path)
    enc = tok(prompts, padding=True, return_tensors="pt")
    out = HuggingFaceGenerationAdapter(model).generate(enc.input_ids, attention_mask=enc.attention_mask, max_new_tokens=a.max_new_tokens)
    if int(os.environ.get("RANK", "0")) == 0:
        for p, o in zip(prompts, tok.batch_decode(out, skip_special_tokens=True)):
            print("-" * 80 + f"\n{o}")


if __name__ == "__main__":
    main()
