"""reference modules/moe_v2.py (``initialize_moe_module`` building router + ExpertMLPsV2 + shared experts) — same factory,
implementation in modules/moe.py."""
from .moe import MoE, ExpertMLPs, RouterTopK, SharedExperts, initialize_moe_module  # noqa: F401

ExpertMLPsV2 = ExpertMLPs
