"""Module-level cross-backend test templates (reference module_test/base_template/{adapter,orchestrator}_base.py and
module_from_model_template/*): run ONE module three ways — a golden implementation (usually Hugging Face) on CPU, the engine's
module on CPU (fp32 oracle path), the engine's module on the GPU (CUDA kernels) — on identical weights and inputs, and compare.

    class RMSNormAdapter(ModuleAdapter):
        def build_golden(self):   return HFRMSNorm(64)
        def build_engine(self, device, dtype):  return RMSNorm(64, dtype=dtype, device=device)
        def transfer_weights(self, golden, engine): engine.weight.copy_(golden.weight)
        def make_inputs(self):    return (torch.randn(2, 5, 64),)
    ModuleTestOrchestrator(RMSNormAdapter()).run()
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn


class ModuleAdapter:
    """Subclass per module under test."""
    rtol, atol = 1e-2, 1e-3
    gpu_dtype = torch.bfloat16

    def build_golden(self) -> nn.Module:
        raise NotImplementedError

    def build_engine(self, device: torch.device, dtype: torch.dtype) -> nn.Module:
        raise NotImplementedError

    def transfer_weights(self, golden: nn.Module, engine: nn.Module) -> None:
        engine.load_state_dict(golden.state_dict(), strict=False)

    def make_inputs(self) -> Tuple[torch.Tensor, ...]:
        raise NotImplementedError

    def call_golden(self, m, *inputs):
        return m(*inputs)

    def call_engine(self, m, *inputs):
        return m(*inputs)


class ModuleFromModelAdapter(ModuleAdapter):
    """Extract the module under test from a full model by path (reference module_from_model_template)."""
    module_path = ""

    def from_model(self, model: nn.Module) -> nn.Module:
        m = model
        for part in self.module_path.split("."):
            m = m[int(part)] if part.isdigit() else getattr(m, part)
        return m


class ModuleTestOrchestrator:
    def __init__(self, adapter: ModuleAdapter):
        self.a = adapter

    @torch.no_grad()
    def run(self, devices=("cpu", "cuda")) -> Dict[str, float]:
        a = self.a
        golden = a.build_golden().eval().float()
        inputs = a.make_inputs()
        exp = a.call_golden(golden, *inputs)
        exp = exp[0] if isinstance(exp, (tuple, list)) else exp
        report = {}
        for dev in devices:
            if dev == "cuda" and not torch.cuda.is_available():
                continue
            dt = torch.float32 if dev == "cpu" else a.gpu_dtype
            eng = a.build_engine(torch.device(dev), dt).eval()
            a.transfer_weights(golden, eng)
            got = a.call_engine(eng, *[x.to(dev, dt if x.is_floating_point() else x.dtype) for x in inputs])
            got = got[0] if isinstance(got, (tuple, list)) else got
            err = (got.float().cpu() - exp.float()).abs().max().item()
            tol = (a.atol + a.rtol * exp.abs().max().item()) * (1 if dev == "cpu" else 4)
            report[dev] = err
            assert err <= tol, f"{type(a).__name__} on {dev}: max abs err {err} > {tol}"
        return report
