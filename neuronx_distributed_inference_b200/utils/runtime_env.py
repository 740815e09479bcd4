"""Runtime environment variables derived from a NeuronConfig (reference utils/runtime_env.py:1-45 sets NEURON_RT_* variables;
the B200 counterparts are NCCL / CUDA knobs).  Only variables not already present in the environment are set."""
from __future__ import annotations

import os
from typing import Dict


def get_env_vars(neuron_config) -> Dict[str, str]:
    env: Dict[str, str] = {}
    if neuron_config.tp_degree > 1:
        env["NCCL_NVLS_ENABLE"] = "1"                 # in-switch reductions on NVSwitch when available
        env["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "1"
        env["CUDA_DEVICE_MAX_CONNECTIONS"] = "8"
    if getattr(neuron_config, "enable_long_context_mode", False):
        env["TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC"] = "600"   # very long prefill steps must not trip the watchdog
    if not getattr(neuron_config, "cuda_graphs", True):
        env["NXDI_B200_PDL"] = os.environ.get("NXDI_B200_PDL", "1")
    return env


def set_env_vars(neuron_config) -> None:
    for k, v in get_env_vars(neuron_config).items():
        os.environ.setdefault(k, str(v))
