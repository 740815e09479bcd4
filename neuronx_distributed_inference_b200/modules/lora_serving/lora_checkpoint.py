"""Adapter checkpoints (reference modules/lora_serving/lora_checkpoint.py:19-412): read PEFT directories (``adapter_config.json`` +
``adapter_model.safetensors``), validate them against the serving config (rank, target modules), and keep the host-side pool that
dynamic multi-LoRA swaps from."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from ..lora import LoraModel, _HF_TO_FUSED


class LoraCheckpoint:
    def __init__(self, lora_config):
        self.lora_config = lora_config
        self.cpu_pool: Dict[str, Tuple[dict, Dict[str, torch.Tensor]]] = {}

    # ---- reading -------------------------------------------------------------------------------------------------------------
    @staticmethod
    def read(path: str) -> Tuple[dict, Dict[str, torch.Tensor]]:
        return LoraModel.read_peft(path)

    def validate(self, name: str, cfg: dict, sd: Dict[str, torch.Tensor]):
        """Rank within ``max_lora_rank``; every adapted module is one of the served ``target_modules``."""
        ranks = {v.shape[0] for k, v in sd.items() if "lora_A" in k}
        if ranks and max(ranks) > self.lora_config.max_lora_rank:
            raise ValueError(f"adapter {name}: rank {max(ranks)} > max_lora_rank {self.lora_config.max_lora_rank}")
        served = set(self.lora_config.target_modules or ["q_proj", "k_proj", "v_proj", "o_proj"])
        used = {t for t in _HF_TO_FUSED if any(f".{t}." in k for k in sd)}
        extra = used - served
        if extra:
            raise ValueError(f"adapter {name} adapts {sorted(extra)} which are not in target_modules {sorted(served)}")
        return cfg.get("lora_alpha", self.lora_config.lora_alpha), (max(ranks) if ranks else 0)

    # ---- host pool (dynamic multi-LoRA) -----------------------------------------------------------------------------------------
    def load_to_cpu(self, name: str, path: Optional[str] = None, state_dict: Optional[dict] = None, cfg: Optional[dict] = None):
        if state_dict is None:
            cfg, state_dict = self.read(path)
        cfg = cfg or {}
        self.validate(name, cfg, state_dict)
        limit = self.lora_config.max_cpu_loras
        if limit and name not in self.cpu_pool and len(self.cpu_pool) >= limit:
            raise RuntimeError(f"host adapter pool is full ({limit}); remove an adapter first")
        self.cpu_pool[name] = (cfg, {k: (v.pin_memory() if torch.cuda.is_available() else v) for k, v in state_dict.items()})
        return self.cpu_pool[name]

    def get(self, name: str):
        return self.cpu_pool[name]

    def remove(self, name: str):
        self.cpu_pool.pop(name, None)

    def load_all_configured(self):
        """Adapters named in ``lora_ckpt_paths`` (device-resident at start) and ``lora_ckpt_paths_cpu`` (host pool)."""
        out = {}
        for name, path in {**self.lora_config.lora_ckpt_paths, **self.lora_config.lora_ckpt_paths_cpu}.items():
            out[name] = self.load_to_cpu(name, path)
        return out
