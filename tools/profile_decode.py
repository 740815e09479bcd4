"""Short decode loop for ncu: N-layer Llama-3.1-8B-shaped model, a prefill and a few eager decode steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LLAMA31_8B, build_app  # noqa: E402

layers = int(os.environ.get("LAYERS", "2"))
steps = int(os.environ.get("STEPS", "3"))
app = build_app(dict(LLAMA31_8B, num_hidden_layers=layers), 1, 2, 256, 128, False)
app.neuron_config.cuda_graphs = False
app.token_generation_model.use_graphs = False
ids = torch.randint(0, 100, (2, 128))
tok = app(ids, attention_mask=torch.ones_like(ids)).tokens
pos = torch.full((2, 1), 128, dtype=torch.int32)
torch.cuda.synchronize()
for _ in range(steps):
    tok = app(tok.view(2, 1), position_ids=pos).tokens
    pos += 1
torch.cuda.synchronize()
print("done", tok.tolist())
