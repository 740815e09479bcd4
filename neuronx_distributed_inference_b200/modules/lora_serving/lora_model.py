"""reference modules/lora_serving/lora_model.py: ``LoraModel`` :36-201, ``LoraWeightManager`` :204-260, ``AdapterCache`` :294-422,
``LoraModelManager`` :425-682."""
from __future__ import annotations

from typing import List

import torch

from ..lora import AdapterCache, LoraModel, LoraModelManager  # noqa: F401
from .lora_checkpoint import LoraCheckpoint


class LoraWeightManager:
    """Bookkeeping around the device-resident LoRA tensors of a served model: enumerate them, report their footprint, and — for
    dynamic multi-LoRA — make sure the adapters a batch asks for are in device slots before the step runs, returning the slot ids
    (the reference swaps whole weight tensors in ``update_lora_tensors``; here a swap is a host-to-device copy into the slot)."""

    def __init__(self, config, base_model: LoraModel = None, manager: LoraModelManager = None):
        self.lora_config = config
        self.base_model = base_model
        self.manager = manager
        self.lora_checkpoint = LoraCheckpoint(config)

    @staticmethod
    def _is_lora_module(name: str) -> bool:
        return name.endswith(".A") or name.endswith(".B") or "lora_A" in name or "lora_B" in name

    def get_lora_tensors(self) -> List[torch.Tensor]:
        if self.base_model is None:
            raise ValueError("Base model is not set for LoraWeightManager.")
        return [p for n, p in self.base_model.named_parameters() if self._is_lora_module(n)]

    def update_lora_adapter_ids(self, adapter_names, device=None) -> torch.Tensor:
        """Adapter NAMES of the rows of a batch -> device slot ids, swapping adapters in from the host pool when needed."""
        if self.manager is None:
            raise ValueError("dynamic multi-LoRA needs a LoraModelManager")
        return self.manager.adapter_ids(list(adapter_names), device)

    def update_lora_tensors(self, adapter_names, device=None) -> torch.Tensor:
        return self.update_lora_adapter_ids(adapter_names, device)

    def lora_memory_footprint(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.get_lora_tensors())

    def print_lora_memory_footprint(self):
        b = self.lora_memory_footprint()
        print(f"LoRA weights on device: {b / 2 ** 20:.1f} MiB in {len(self.get_lora_tensors())} tensors "
              f"({self.lora_config.max_loras} slots, rank <= {self.lora_config.max_lora_rank})")
        return b
