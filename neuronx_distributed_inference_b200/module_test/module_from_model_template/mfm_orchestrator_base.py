"""reference module_test/module_from_model_template/mfm_orchestrator_base.py:17-34."""
from __future__ import annotations

from dataclasses import dataclass

from ..base_template.orchestrator_base import OrchestratorBase, OrchestratorBaseConfig


@dataclass(kw_only=True)
class MFMOrchestratorConfig(OrchestratorBaseConfig):
    layer_id: int = 0
    model_tag: str = "context_encoding_model"      # which sub-model of the application the modules are cut from (one model on B200)


class MFMOrchestratorBase(OrchestratorBase):
    """Random hidden states / KV cache of the base orchestrator; kept as the extension point for recorded inputs."""
