"""Seeding (reference utils/random.py)."""
import random as _random

import torch


def set_random_seed(seed: int = 0):
    _random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    try:
        import numpy as np
        np.random.seed(seed)
    except Exception:  # pragma: no cover
        pass
