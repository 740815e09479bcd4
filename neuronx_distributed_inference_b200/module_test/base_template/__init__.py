from .adapter_base import (B200DeviceAdapterBase, HFAdapterBase, ModuleAdapterBase, NxDINeuronAdapterBase,  # noqa: F401
                           NxDISingleRankCPUAdapterBase)
from .orchestrator_base import OrchestratorBase, OrchestratorBaseConfig  # noqa: F401
