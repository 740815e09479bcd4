"""reference experimental/core/config/build.py:5 — knobs of a build flow.  ``compiler_args`` is kept for YAML compatibility (nothing is
compiled on B200); ``cuda_graphs`` decides whether the built model replays its per-tag forward from a CUDA graph."""
from dataclasses import dataclass, field
from typing import List


@dataclass
class BuildConfig:
    world_size: int = 1
    batch_size: int = 1
    sequence_length: int = 1024
    sequence_length_buckets: dict = field(default_factory=dict)
    compiler_args: List[str] = field(default_factory=list)
    cuda_graphs: bool = True
