"""Context-parallel process-group meshes (reference experimental/functional/pg/context_parallel.py:28-207).

``world_size`` ranks form ``cp_degree`` tensor-parallel groups of ``world_size / cp_degree`` CONTIGUOUS ranks (attention runs
tensor-parallel inside one of them on its slice of the sequence) and ``world_size / cp_degree`` context-parallel groups of ranks with the
same index inside their TP group (K/V are gathered across them)."""
from __future__ import annotations

from typing import List

import torch

from ....parallel import state as pstate


def get_context_parallel_tp_mesh(world_size: int, cp_degree: int) -> List[List[int]]:
    _check(world_size, cp_degree)
    tp = world_size // cp_degree
    return [list(range(g * tp, (g + 1) * tp)) for g in range(cp_degree)]


def get_context_parallel_cp_mesh(world_size: int, cp_degree: int) -> List[List[int]]:
    _check(world_size, cp_degree)
    tp = world_size // cp_degree
    return [[i + g * tp for g in range(cp_degree)] for i in range(tp)]


def get_cp_rank(rank: torch.Tensor, world_size: int, cp_degree: int) -> torch.Tensor:
    """Index of the TP group (= sequence slice) a global rank belongs to."""
    _check(world_size, cp_degree)
    return torch.div(torch.as_tensor(rank), world_size // cp_degree, rounding_mode="floor").to(torch.int32)


def initialize_context_parallel_process_groups(world_size: int, cp_degree: int) -> None:
    pstate.initialize_model_parallel(tensor_model_parallel_size=world_size, context_parallel_size=cp_degree)


initialize_context_parallel_tp_group = initialize_context_parallel_cp_group = initialize_context_parallel_process_groups


def get_context_parallel_tp_group():
    return pstate.get_context_parallel_tp_group()


def get_context_parallel_cp_group():
    return pstate.get_context_parallel_group()


def _check(world_size: int, cp_degree: int):
    if cp_degree < 1 or world_size % cp_degree != 0:
        raise ValueError(f"cp_degree {cp_degree} must divide world_size {world_size}")
