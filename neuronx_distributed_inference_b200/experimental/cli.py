"""Experimental CLI (reference experimental/cli.py:1-281): build a functional model from a YAML config + checkpoint and run greedy
generation.   python -m neuronx_distributed_inference_b200.experimental.cli --config cfg.yaml --model-path <hf dir> --prompt-ids 1,2,3"""
from __future__ import annotations

import argparse

import torch


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--config", required=True)
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--prompt-ids", default="1,2,3,4")
    ap.add_argument("--max-new-tokens", type=int, default=16)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    a = ap.parse_args(argv)
    from ..config import load_pretrained_config
    from ..models.llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaForCausalLM
    from .core import NeuronConfigHandler, generate, load_yaml_config
    from .models.llama3.model import Llama3, Llama3Args
    h = NeuronConfigHandler(load_yaml_config(a.config))
    nc = h.common_config()
    cfg = LlamaInferenceConfig(nc, load_config=load_pretrained_config(a.model_path))
    sd = NeuronLlamaForCausalLM.get_state_dict(a.model_path, cfg)
    args = Llama3Args(dim=cfg.hidden_size, n_layers=cfg.num_hidden_layers, n_heads=cfg.num_attention_heads,
                      n_kv_heads=cfg.num_key_value_heads, vocab_size=cfg.vocab_size, ffn_dim=cfg.intermediate_size,
                      norm_eps=cfg.rms_norm_eps, rope_theta=getattr(cfg, "rope_theta", 10000.0) or 10000.0,
                      rope_scaling=getattr(cfg, "rope_scaling", None), max_batch_size=nc.batch_size, max_seq_len=nc.seq_len,
                      dtype=nc.torch_dtype)
    model = Llama3(args, sd, torch.device(a.device))
    ids = torch.tensor([[int(x) for x in a.prompt_ids.split(",")]])
    print(generate(model, ids, max_new_tokens=a.max_new_tokens).tolist())


if __name__ == "__main__":
    main()
