"""Text generation with a Hugging Face Llama checkpoint (role of the reference's examples/generation_demo.py).

    python examples/generation_demo.py --model-path /path/to/Llama-3.1-8B-Instruct --tp-degree 1 --prompt "I believe the meaning of life is"
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/generation_demo.py --model-path ... --tp-degree 8
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import argparse

import torch

from neuronx_distributed_inference_b200.config import NeuronConfig, OnDeviceSamplingConfig, load_pretrained_config
from neuronx_distributed_inference_b200.models.llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaForCausalLM
from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--compiled-path", default="/tmp/nxdi_b200_artifacts")
    ap.add_argument("--tp-degree", type=int, default=1)
    ap.add_argument("--prompt", action="append")
    ap.add_argument("--max-new-tokens", type=int, default=64)
    ap.add_argument("--top-k", type=int, default=1)
    a = ap.parse_args()
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(a.model_path, padding_side="right")
    tok.pad_token = tok.pad_token or tok.eos_token
    prompts = a.prompt or ["I believe the meaning of life is", "The color of the sky is"]
    nc = NeuronConfig(tp_degree=a.tp_degree, batch_size=len(prompts), max_context_length=256, seq_len=512, torch_dtype="bfloat16",
                      on_device_sampling_config=OnDeviceSamplingConfig(top_k=a.top_k, do_sample=a.top_k > 1, dynamic=a.top_k > 1),
                      enable_bucketing=True)
    cfg = LlamaInferenceConfig(nc, load_config=load_pretrained_config(a.model_path))
    model = NeuronLlamaForCausalLM(a.model_path, cfg)
    model.compile(a.compiled_path)        # writes the config (+ pre-sharded weights with save_sharded_checkpoint)
    model.load(a.compiled_path)
    enc = tok(prompts, padding=True, return_tensors="pt")
    out = HuggingFaceGenerationAdapter(model).generate(enc.input_ids, attention_mask=enc.attention_mask,
                                                       max_new_tokens=a.max_new_tokens, eos_token_id=tok.eos_token_id,
                                                       pad_token_id=tok.pad_token_id)
    for i, text in enumerate(tok.batch_decode(out, skip_special_tokens=True)):
        print(f"--- output {i}\n{text}")


if __name__ == "__main__":
    main()
