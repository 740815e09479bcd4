"""Input capture hook (``--capture-indices``; reference utils/debug_utils.py:11-97): save the forward() inputs of the N-th
calls of the application to ``<dir>/saved_inputs_<i>.pt``."""
from __future__ import annotations

import os
from typing import Iterable

import torch


def capture_model_inputs(app, capture_indices: Iterable[int], save_dir: str = "saved_inputs"):
    idx = set(capture_indices)
    state = {"i": 0}
    orig = app.forward

    def wrapped(*a, **k):
        if state["i"] in idx:
            os.makedirs(save_dir, exist_ok=True)
            blob = {"args": [x.detach().cpu() if torch.is_tensor(x) else x for x in a],
                    "kwargs": {n: (v.detach().cpu() if torch.is_tensor(v) else v) for n, v in k.items()}}
            torch.save(blob, os.path.join(save_dir, f"saved_inputs_{state['i']}.pt"))
        state["i"] += 1
        return orig(*a, **k)
    app.forward = wrapped
    return app
