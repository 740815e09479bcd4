"""Backend adapters of the module-test template (reference module_test/base_template/adapter_base.py:31-205).

An adapter runs ONE module on ONE backend in four steps — ``define_module_cls`` (what the module is and how it is called),
``instantiate_module`` (build it and load its weights), ``load_kv_cache`` (optional: a Hugging Face layout cache ``(k, v)`` of
``[B, H_kv, S, D]``), ``run_inference`` — so that an orchestrator can feed identical inputs to a Hugging Face golden, the engine on the CPU
(fp32 PyTorch definitions) and the engine on the GPU (CUDA kernels) and compare the outputs pairwise.

Backends here: :class:`HFAdapterBase`, :class:`NxDISingleRankCPUAdapterBase` (engine, CPU), :class:`B200DeviceAdapterBase` (engine on
``cuda:0``; ``NxDINeuronAdapterBase`` is kept as an alias of it — there is nothing to trace or compile on B200, "instantiate" = build
on the device, optionally with the tensor-parallel world already initialised by the launcher)."""
from __future__ import annotations

import gc
import logging
from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Type

import torch
import torch.nn as nn


class ModuleAdapterBase(ABC):
    def __init__(self):
        self.module_cls: Optional[Type[nn.Module]] = None
        self.module: Optional[nn.Module] = None
        self.logger = logging.getLogger("b200infer")
        self.torch_dtype: torch.dtype = torch.float32
        self.hf_ckpt_path: Optional[str] = None
        self.device = torch.device("cpu")

    # ---- the four steps -----------------------------------------------------------------------------------------------------
    @abstractmethod
    def define_module_cls(self, *args, **kwargs):
        """Set ``self.module_cls`` (a zero-argument ``nn.Module`` factory whose ``forward`` is the call under test)."""

    def instantiate_module(self, *args, **kwargs):
        """Build ``self.module`` from ``self.module_cls`` and load ``self.get_state_dict()`` into it."""
        assert self.module_cls is not None, "define_module_cls() first"
        m = self.module_cls()
        sd = self.get_state_dict()
        if sd:
            missing, unexpected = m.load_state_dict(sd, strict=False)
            if unexpected:
                raise KeyError(f"{type(self).__name__}: weights without a parameter: {list(unexpected)[:6]}")
            if missing:
                self.logger.warning("%s: parameters without weights keep their initial values: %s", type(self).__name__, list(missing)[:6])
        self.module = m.to(device=self.device, dtype=self.torch_dtype).eval()

    def load_kv_cache(self, hf_kv_cache: Any):
        raise NotImplementedError(f"{type(self).__name__}.load_kv_cache() is not implemented")

    def run_inference(self, *args, **kwargs):
        assert self.module is not None, "instantiate_module() first"
        move = lambda t: t.to(self.device, self.torch_dtype if t.is_floating_point() else t.dtype) if torch.is_tensor(t) else t  # noqa: E731
        with torch.no_grad():
            out = self.module(*[move(a) for a in args], **{k: move(v) for k, v in kwargs.items()})
        out = out[0] if isinstance(out, (tuple, list)) else out
        return out.float().cpu()

    # ---- helpers ------------------------------------------------------------------------------------------------------------
    def get_state_dict(self) -> Dict[str, torch.Tensor]:
        return {}

    def set_torch_dtype(self, torch_dtype: torch.dtype):
        self.torch_dtype = torch_dtype

    def set_hf_ckpt_path(self, hf_ckpt_path: str):
        self.hf_ckpt_path = hf_ckpt_path

    def free_resources(self):
        self.module = None
        self.module_cls = None
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()


class HFAdapterBase(ModuleAdapterBase):
    """Golden backend: a Hugging Face module, strict weight loading, always on the CPU."""

    def instantiate_module(self):
        assert self.module_cls is not None
        m = self.module_cls()
        sd = self.get_state_dict()
        if sd:
            m.load_state_dict(sd)
        self.module = m.to(self.torch_dtype).eval()


class NxDISingleRankCPUAdapterBase(ModuleAdapterBase):
    """The engine's module on the CPU, single rank (the fp32 PyTorch definitions of every op)."""

    def instantiate_module(self):
        from ...utils.testing import init_cpu_env
        init_cpu_env(1)
        super().instantiate_module()


class B200DeviceAdapterBase(ModuleAdapterBase):
    """The engine's module on the GPU through the hand-written kernels.  ``tp_degree`` / ``world_size`` describe the launch the
    test runs under (torchrun for > 1; the parallel state must then be initialised by the worker)."""

    def __init__(self, tp_degree: int = 1, world_size: int = 1):
        super().__init__()
        self.tp, self.ws = tp_degree, world_size
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    def instantiate_module(self, example_inputs=None):
        if self.device.type != "cuda":
            raise RuntimeError("the device adapter needs a GPU (skip it on CPU-only hosts)")
        from ...parallel import state as pstate
        if self.tp > 1 and (not pstate.model_parallel_is_initialized() or pstate.get_tensor_model_parallel_size() != self.tp):
            raise RuntimeError(f"tp_degree={self.tp}: run under torchrun and initialise the model-parallel state first")
        super().instantiate_module()


NxDINeuronAdapterBase = B200DeviceAdapterBase      # the reference's name for "the accelerator backend"
