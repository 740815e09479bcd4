"""Orchestrator of the module-test template (reference module_test/base_template/orchestrator_base.py:26-218): one set of random inputs
(and, optionally, one random KV cache in the Hugging Face layout) for every adapter, adapters processed one after the other (each
frees its module before the next is built), outputs compared pairwise."""
from __future__ import annotations

import logging
from dataclasses import dataclass
from itertools import combinations
from typing import Dict, List, Optional, Tuple

import torch

from ...utils.accuracy import check_accuracy_embeddings
from .adapter_base import B200DeviceAdapterBase, ModuleAdapterBase


@dataclass(kw_only=True)
class OrchestratorBaseConfig:
    batch_size: int
    torch_dtype: torch.dtype
    hf_weight_ckpt_path: Optional[str] = None
    atol: Optional[float] = None
    rtol: Optional[float] = None
    seed: int = 0

    @dataclass
    class PrepInputConfig:
        seq_len: int                       # 1 for token generation, the context length for context encoding
        hf_hidden_size: int
        hidden_states_mean: float = 0.0
        hidden_states_std: float = 1.0

    @dataclass
    class PrepKVCacheConfig:
        ctx_len: int
        num_head: int
        hf_head_hidden_size: int
        kv_cache_mean: float = 0.0
        kv_cache_std: float = 1.0

    prep_input_config: "OrchestratorBaseConfig.PrepInputConfig"
    prep_kv_cache_config: Optional["OrchestratorBaseConfig.PrepKVCacheConfig"] = None


class OrchestratorBase:
    DEFAULT_ATOL, DEFAULT_RTOL = 1e-3, 1e-2

    def __init__(self, adapters: List[ModuleAdapterBase], orchestratorConf: OrchestratorBaseConfig):
        self.adapters, self.conf = adapters, orchestratorConf
        self.logger = logging.getLogger("b200infer")
        for a in adapters:
            a.set_hf_ckpt_path(orchestratorConf.hf_weight_ckpt_path)
            if not isinstance(a, B200DeviceAdapterBase):     # the device adapter keeps its kernel dtype unless told otherwise
                a.set_torch_dtype(orchestratorConf.torch_dtype)

    def prepare_input_hf_format(self) -> Dict[str, torch.Tensor]:
        c, p = self.conf, self.conf.prep_input_config
        g = torch.Generator().manual_seed(c.seed)
        x = torch.randn(c.batch_size, p.seq_len, p.hf_hidden_size, generator=g) * p.hidden_states_std + p.hidden_states_mean
        return {"hidden_states": x.to(c.torch_dtype)}

    def prepare_kv_cache_hf_format(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(k, v), each ``[batch, num_head, ctx_len, head_dim]``."""
        c, p = self.conf, self.conf.prep_kv_cache_config
        g = torch.Generator().manual_seed(c.seed + 1)
        shape = (c.batch_size, p.num_head, p.ctx_len, p.hf_head_hidden_size)
        mk = lambda: (torch.randn(shape, generator=g) * p.kv_cache_std + p.kv_cache_mean).to(c.torch_dtype)      # noqa: E731
        return mk(), mk()

    @classmethod
    def validate_result(cls, adp_name_to_result: Dict[str, torch.Tensor], atol: Optional[float] = None, rtol: Optional[float] = None):
        atol = cls.DEFAULT_ATOL if atol is None else atol
        rtol = cls.DEFAULT_RTOL if rtol is None else rtol
        errs = {}
        for a, b in combinations(list(adp_name_to_result), 2):
            ok, err = check_accuracy_embeddings(adp_name_to_result[a], adp_name_to_result[b], rtol=rtol, atol=atol)
            errs[(a, b)] = err
            if not ok:
                raise AssertionError(f"[mismatch] {a} vs {b}: max abs error {err} (atol {atol}, rtol {rtol})")
        return errs

    def run_validation(self):
        inputs = self.prepare_input_hf_format()
        kv = self.prepare_kv_cache_hf_format() if self.conf.prep_kv_cache_config else None
        results: Dict[str, torch.Tensor] = {}
        for i, adp in enumerate(self.adapters, 1):
            name = type(adp).__name__
            self.logger.info("[%s] adapter %d/%d: %s", type(self).__name__, i, len(self.adapters), name)
            adp.define_module_cls()
            if isinstance(adp, B200DeviceAdapterBase):
                adp.instantiate_module(tuple(inputs.values()))
            else:
                adp.instantiate_module()
            if kv is not None:
                adp.load_kv_cache(kv)
            results[name] = adp.run_inference(**inputs)
            adp.free_resources()
        return self.validate_result(results, atol=self.conf.atol, rtol=self.conf.rtol)
