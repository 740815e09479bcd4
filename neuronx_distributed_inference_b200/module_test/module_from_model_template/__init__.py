from .mfm_adapter_base import (MFMB200DeviceAdapter, MFMHFAdapter, MFMNxDICPUSingleRankAdapter, MFMNxDINeuronAdapter,  # noqa: F401
                               build_prefixes_map, extract_submodules_by_prefixes, extract_subweights_by_prefixes)
from .mfm_orchestrator_base import MFMOrchestratorBase, MFMOrchestratorConfig  # noqa: F401
