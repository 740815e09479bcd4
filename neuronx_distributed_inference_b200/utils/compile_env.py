"""Build-time environment (reference utils/compile_env.py:1-45 exports compiler variables before tracing).  Nothing is traced on
B200; "compile time" is the nvcc build of the extension, steered by these variables (read by ops/_ext.py)."""
from __future__ import annotations

import os
from typing import Dict


def get_compile_env_vars(neuron_config) -> Dict[str, str]:
    env = {"NXDI_B200_ARCH": "sm_100a"}
    if getattr(neuron_config, "logical_nc_config", None):
        pass   # Neuron logical-core setting: no analogue
    return env


def set_compile_env_vars(neuron_config) -> None:
    for k, v in get_compile_env_vars(neuron_config).items():
        os.environ.setdefault(k, str(v))
