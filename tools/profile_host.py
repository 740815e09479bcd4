"""cProfile of the host side of one decode step through the public forward() API."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LLAMA31_8B, build_app  # noqa: E402

app = build_app(dict(LLAMA31_8B, num_hidden_layers=int(os.environ.get("LAYERS", "4"))), 1, 2, 272, 128, False)
ids = torch.randint(0, 100, (2, 128))
tok = app(ids, attention_mask=torch.ones_like(ids)).tokens.cpu()
pos = torch.full((2, 1), 128, dtype=torch.int32)
for _ in range(5):
    tok = app(tok.view(2, 1), position_ids=pos).tokens.cpu(); pos += 1
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    tok = app(tok.view(2, 1), position_ids=pos).tokens.cpu(); pos += 1
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
