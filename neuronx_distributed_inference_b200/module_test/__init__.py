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


# ---- module-from-model template with shared inputs and a pre-filled KV cache --------------------------------------------------
# reference module_test/base_template/orchestrator_base.py (``prepare_input_hf_format`` / ``prepare_kv_cache_hf_format`` /
# ``run_validation``) and module_from_model_template/mfm_adapter_base.py (HF / NxDI-CPU / NxDI-device adapters of ONE decoder layer
# cut out of a full model).
from dataclasses import dataclass  # noqa: E402


@dataclass
class OrchestratorConfig:
    batch_size: int = 2
    seq_len: int = 1                  # tokens fed to the module (1 = a decode step)
    past_len: int = 6                 # tokens already in the KV cache (0 = prefill)
    layer_idx: int = 0
    seed: int = 0
    rtol: float = 1e-2
    atol: float = 1e-3


class DecoderLayerFromModelOrchestrator:
    """Runs decoder layer ``layer_idx`` of (a) a Hugging Face causal LM and (b) the engine application built from the SAME checkpoint,
    on identical random hidden states and an identical random KV cache (given in the Hugging Face layout ``[B, H_kv, S, D]`` and loaded
    into the engine's cache lines), and compares the layer outputs — on CPU (fp32) and, when present, on the GPU kernels."""

    def __init__(self, hf_model, app_factory, config: OrchestratorConfig = OrchestratorConfig()):
        self.hf, self.app_factory, self.c = hf_model.eval(), app_factory, config

    # shared inputs ------------------------------------------------------------------------------------------------------------
    def prepare_input_hf_format(self):
        c, hc = self.c, self.hf.config
        g = torch.Generator().manual_seed(c.seed)
        hidden = torch.randn(c.batch_size, c.seq_len, hc.hidden_size, generator=g)
        pos = (torch.arange(c.seq_len) + c.past_len).unsqueeze(0).expand(c.batch_size, -1)
        return dict(hidden_states=hidden, position_ids=pos)

    def prepare_kv_cache_hf_format(self):
        c, hc = self.c, self.hf.config
        g = torch.Generator().manual_seed(c.seed + 1)
        D = getattr(hc, "head_dim", None) or hc.hidden_size // hc.num_attention_heads
        shape = (c.batch_size, hc.num_key_value_heads, c.past_len, D)
        return torch.randn(shape, generator=g), torch.randn(shape, generator=g)

    # backends ------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run_hf(self, inputs, kv):
        from transformers import DynamicCache
        c = self.c
        layer = self.hf.model.layers[c.layer_idx]
        cache = DynamicCache(config=self.hf.config)
        if c.past_len:
            for i in range(c.layer_idx + 1):        # earlier layers get the same (unused) content so that the cache length is consistent
                cache.update(kv[0].clone(), kv[1].clone(), i)
        pos = inputs["position_ids"]
        emb = self.hf.model.rotary_emb(inputs["hidden_states"], pos)
        T, S = c.seq_len, c.past_len + c.seq_len
        mask = torch.full((T, S), float("-inf")).triu(c.past_len + 1)[None, None].expand(c.batch_size, 1, T, S)
        out = layer(inputs["hidden_states"], attention_mask=mask, position_ids=pos, past_key_values=cache, position_embeddings=emb)
        return out[0] if isinstance(out, (tuple, list)) else out

    @torch.no_grad()
    def run_engine(self, inputs, kv, device: str):
        from ..modules.attention.attention_base import AttnMeta
        c = self.c
        app = self.app_factory(device)
        model = app.model
        dt = next(model.parameters()).dtype
        mgr = model.kv_mgr
        mgr.reset()
        k_cache, v_cache = mgr.get_kv_by_layer_id(c.layer_idx)
        if c.past_len:                                   # HF layout [B, H_kv, S, D] -> cache lines 0..B-1 (this rank's heads at tp=1)
            k_cache[: c.batch_size, :, : c.past_len] = kv[0].to(k_cache.device, k_cache.dtype)
            v_cache[: c.batch_size, :, : c.past_len] = kv[1].to(v_cache.device, v_cache.dtype)
        pos = inputs["position_ids"].to(device).int()
        meta = AttnMeta(is_prefill=c.past_len == 0, position_ids=pos, write_positions=pos, seq_ids=torch.arange(c.batch_size, device=device).int(),
                        seq_hint=c.past_len + c.seq_len)
        return model.layers[c.layer_idx](inputs["hidden_states"].to(device, dt), meta, mgr).float().cpu()

    def validate_result(self, name, got, exp):
        err = (got - exp).abs().max().item()
        tol = self.c.atol + self.c.rtol * exp.abs().max().item()
        assert err <= tol, f"{name}: max abs err {err} > {tol}"
        return err

    def run_validation(self, devices=("cpu", "cuda")):
        inputs, kv = self.prepare_input_hf_format(), self.prepare_kv_cache_hf_format()
        exp = self.run_hf(inputs, kv).float()
        report = {}
        for dev in devices:
            if dev == "cuda" and not torch.cuda.is_available():
                continue
            report[dev] = self.validate_result(dev, self.run_engine(inputs, kv, dev), exp)
        return report
