"""reference layout modules/flashdecode/utils.py"""
from .utils import calculate_num_cores_per_group, combine, local_horizon, local_slots, partial_attention  # noqa: F401
