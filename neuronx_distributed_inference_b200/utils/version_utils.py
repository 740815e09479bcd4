"""Version checks (reference utils/version_utils.py gates features on the installed neuronx-cc / torch-neuronx)."""
from __future__ import annotations

import torch


def get_torch_version():
    return tuple(int(x) for x in torch.__version__.split("+")[0].split(".")[:2])


def get_cuda_version():
    return torch.version.cuda


def get_device_capability():
    return torch.cuda.get_device_capability() if torch.cuda.is_available() else None


def is_blackwell() -> bool:
    cap = get_device_capability()
    return cap is not None and cap[0] == 10
