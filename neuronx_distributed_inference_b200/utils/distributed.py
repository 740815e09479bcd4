"""Rank helpers (reference utils/distributed.py): which ranks live in this process / node, rank-0 gating, barriers."""
from __future__ import annotations

import os

import torch.distributed as dist


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else int(os.environ.get("RANK", "0"))


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else int(os.environ.get("WORLD_SIZE", "1"))


def get_local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def is_rank_zero() -> bool:
    return get_rank() == 0


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def get_init_world_size(neuron_config) -> int:
    return neuron_config.world_size if getattr(neuron_config, "world_size", None) else neuron_config.tp_degree


def get_init_rank(neuron_config) -> int:
    return getattr(neuron_config, "start_rank_id", 0) or 0
