"""Speculative decoding variants on one target (reference: examples with --draft-model-path / --enable-eagle-speculation /
--medusa-tree-json).  Random weights keep the example self-contained; swap ``build_random_llama`` for real checkpoints."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout

import torch

from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter
from neuronx_distributed_inference_b200.utils.testing import build_random_llama

ARCH = dict(hidden_size=512, intermediate_size=1024, num_hidden_layers=4, num_attention_heads=8, num_key_value_heads=4, vocab_size=2048,
            head_dim=64)
dev = "cuda" if torch.cuda.is_available() else "cpu"
dt = "bfloat16" if dev == "cuda" else "float32"
ids = torch.randint(1, 2048, (2, 12))
mask = torch.ones_like(ids)
common = dict(batch_size=2, seq_len=256, max_context_length=32, device=dev, dtype=dt, seed=1)

variants = {
    "fused (draft inside the application)": dict(speculation_length=4, enable_fused_speculation=True,
                                                 fused_draft=dict(hf=dict(num_hidden_layers=1))),
    "EAGLE": dict(speculation_length=4, enable_fused_speculation=True, enable_eagle_speculation=True,
                  fused_draft=dict(hf=dict(num_hidden_layers=1), neuron=dict(is_eagle_draft=True))),
    "EAGLE + token tree": dict(speculation_length=4, enable_fused_speculation=True, enable_eagle_speculation=True,
                               token_tree_config={"0": ["1", "2"], "1": ["3", "4"], "2": ["5"], "3": ["6"]},
                               fused_draft=dict(hf=dict(num_hidden_layers=1), neuron=dict(is_eagle_draft=True))),
    "Medusa": dict(is_medusa=True, num_medusa_heads=3, medusa_speculation_length=8, output_logits=True,
                   medusa_tree=[[0], [1], [0, 0], [0, 1], [1, 0], [0, 0, 0]]),
}
ref = HuggingFaceGenerationAdapter(build_random_llama(ARCH, **common)).generate(ids, attention_mask=mask, max_new_tokens=32)
for name, kw in variants.items():
    app = build_random_llama(ARCH, **common, **kw)
    out = HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=32, return_dict_in_generate=True)
    n = min(out.sequences.shape[1], ref.shape[1])
    same = bool((out.sequences[:, :n] == ref[:, :n]).all())
    st = out.speculation_stats
    print(f"{name:40s} steps={st['steps']:3d} accepted/step={st['accepted'] / max(st['steps'], 1) / 2:.2f} identical_to_greedy={same}")
