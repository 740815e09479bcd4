"""Driver of the persistent decode-step kernel (csrc/decode_step.cu): every decoder layer of a decode step in ONE launch.

A :class:`DecodeStepPlan` is built once per (model, batch, tokens, sequence bucket): it owns the activation buffers of the
step (residual stream ping-pong, qkv, attention output, MLP activation) and the device-side phase table

    per layer:  qkv = W_qkv · rmsnorm(h)      ->  attention (q/k norm, RoPE, cache append, split-KV flash decode)
                h1  = W_o · attn (+all-reduce) + h      ->  u = swiglu(W_gu · rmsnorm(h1))      ->  h = W_d · u (+all-reduce) + h1

Eligibility (`eligible`) mirrors the conditions of the per-kernel fast path (modules/attention/attention_base.py) for dense
Llama-style blocks; everything else keeps the layer-by-layer path.  Reference: the reference fuses within a block (SURVEY §2.3
K2-K5: attention_block_tkg, fused QKV, MLP kernels); on B200 the whole step is one resident kernel.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .. import ops

# Opt-in: correct (tests/test_features_gpu.py) but measured SLOWER than the per-kernel path on B200 so far (Llama-3.1-8B bs 2:
# 4.27 vs 2.82 ms/step at TP1, 1.57 vs 1.15 ms on TP8 shard shapes — profiles/decode_r2.md has the phase timeline and the analysis).
_ENABLED = os.environ.get("NXDI_B200_DECODE_STEP", "0") == "1"


def eligible(model, h: torch.Tensor, meta, kw) -> bool:
    """Decided once per (token count, dtype) for the model's static properties, then per call for the dynamic ones."""
    if not _ENABLED or not h.is_cuda or h.dtype != torch.bfloat16:
        return False
    M = h.shape[0] * h.shape[1]
    cache = model.__dict__.setdefault("_dstep_static_ok", {})
    key = (M, h.shape[1])
    if key not in cache:
        cache[key] = _static_ok(model, h, M)
    if not cache[key]:
        return False
    return (meta.active_mask is None and meta.slot_mapping is None and meta.capture is None and meta.adapter_ids is None
            and meta.rotary_position_ids is None and not meta.has_prefix and not kw.get("deepstack_embeds"))


def _static_ok(model, h, M) -> bool:
    from ..models.model_base import DecoderLayer
    from ..modules.kvcache import KVCacheManager
    from ..modules.mlp import GatedMLP
    nc = model.neuron_config
    kv = model.kv_mgr
    if M > ops.GEMV_MAX_TOKENS or type(kv) is not KVCacheManager or getattr(kv, "k_scale", None) is not None:
        return False
    g = model.tp_group
    if g.size > 1 and (g.symm is None or model.hidden_size > g.symm.n_max):
        return False
    rot = None
    C = ops._C()
    for layer in model.layers:
        attn, mlp = getattr(layer, "self_attn", None), getattr(layer, "mlp", None)
        if type(layer) is not DecoderLayer or not isinstance(mlp, GatedMLP) or attn is None or not hasattr(attn, "chain_eligible"):
            return False
        if not attn.chain_eligible(kv, h.dtype) or attn.dp_group is not None or attn.cp_group is not None:
            return False
        if h.shape[1] * (attn.n_q // attn.n_kv) > 64:
            return False
        rot = rot or attn.rotary_emb
        if attn.rotary_emb is not rot:
            return False
        n1, n2 = layer.input_layernorm, layer.post_attention_layernorm
        if getattr(n1, "weight", None) is None or getattr(n2, "weight", None) is None:
            return False
        lin = [attn.qkv_proj, attn.o_proj, mlp.gate_up_proj, mlp.down_proj]
        for m in lin:
            w = m.weight
            if (w.dtype != torch.bfloat16 or w.dim() != 2 or not w.is_contiguous() or w.shape[1] % 64 != 0
                    or getattr(m, "scale", None) is not None or getattr(m, "sequence_parallel_enabled", False)
                    or not C.gemv2_supported(M, w.shape[1])):
                return False
        for m in (attn.o_proj, mlp.down_proj):
            if getattr(m, "reduce_dtype", None) not in (None, torch.float32) or not getattr(m, "reduce_output", True) \
                    or not getattr(m, "input_is_parallel", True):
                return False
        if mlp.act not in ops._ACT_CODES:
            return False
    return rot is not None


class DecodeStepPlan:
    def __init__(self, model, B: int, T: int, seq_hint: int, device):
        C = ops._C()
        self.C, self.model = C, model
        M = B * T
        g = model.tp_group
        self.symm = g.symm if g.size > 1 else None
        dt = torch.bfloat16
        H = model.hidden_size
        l0 = model.layers[0]
        a0 = l0.self_attn
        self.h_a = torch.empty(M, H, dtype=dt, device=device)
        self.h_b = torch.empty(M, H, dtype=dt, device=device)
        self.qkv = torch.empty(M, (a0.n_q + 2 * a0.n_kv) * a0.head_dim, dtype=dt, device=device)
        self.attn = torch.empty(M, a0.n_q * a0.head_dim, dtype=dt, device=device)
        self.u = torch.empty(M, l0.mlp.down_proj.weight.shape[1], dtype=dt, device=device)
        self.handle = C.dstep_new(M)   # rows of every activation = (b, t) pairs; attention sees them as [B, T]
        if self.symm is not None:
            C.dstep_set_symm(self.handle, self.symm.recv_ptrs, self.symm.step_t, self.symm.rank, self.symm.n_max)
        ar = self.symm is not None
        for layer in model.layers:
            attn, mlp, n1, n2 = layer.self_attn, layer.mlp, layer.input_layernorm, layer.post_attention_layernorm
            k_cache, v_cache = model.kv_mgr.get_kv_by_layer_id(attn.layer_idx)
            qn = attn.q_layernorm.weight if attn.qk_norm == "rms_pre_rope" else None
            kn = attn.k_layernorm.weight if attn.qk_norm == "rms_pre_rope" else None
            C.dstep_add_gemv(self.handle, attn.qkv_proj.weight, self.h_a, attn.qkv_proj.bias, n1.weight, float(n1.variance_epsilon),
                             float(n1.offset), 0, None, self.qkv, False)
            C.dstep_add_attn(self.handle, self.qkv, self.attn, k_cache, v_cache, qn, kn, float(attn.qk_norm_eps), B, T, attn.n_q,
                             attn.n_kv, attn.head_dim, float(attn.scale), int(attn.sliding_window or 0), attn.sinks, int(seq_hint))
            # row-parallel bias: replicated, added once after the reduction (rank 0 only when the reduction is skipped)
            C.dstep_add_gemv(self.handle, attn.o_proj.weight, self.attn, attn.o_proj.bias, None, 0.0, 0.0, 0, self.h_a, self.h_b, ar)
            C.dstep_add_gemv(self.handle, mlp.gate_up_proj.weight, self.h_b, mlp.gate_up_proj.bias, n2.weight,
                             float(n2.variance_epsilon), float(n2.offset), ops._ACT_CODES[mlp.act], None, self.u, False)
            C.dstep_add_gemv(self.handle, mlp.down_proj.weight, self.u, mlp.down_proj.bias, None, 0.0, 0.0, 0, self.h_b, self.h_a, ar)
        C.dstep_finalize(self.handle)
        self.n_ar = int(C.dstep_num_allreduce(self.handle))
        self.B, self.T = B, T

    def run(self, h: torch.Tensor, meta, cos: torch.Tensor, sin: torch.Tensor, lines: torch.Tensor) -> torch.Tensor:
        B, T, H = h.shape
        self.h_a.copy_(h.reshape(B * T, H))
        s = self.symm
        call_base, parity_base = (s.call, s.parity) if s is not None else (0, 0)
        ops.stats["decode_step"] += 1
        self.C.dstep_launch(self.handle, meta.position_ids.contiguous(), meta.write_positions.contiguous(), lines.contiguous(),
                            cos.contiguous(), sin.contiguous(), call_base, parity_base)
        if s is not None:
            s.call += self.n_ar
            s.calls += self.n_ar
            s.parity ^= self.n_ar & 1
        return self.h_a.view(B, T, H)

    def __del__(self):
        try:
            self.C.dstep_free(self.handle)
        except Exception:
            pass


def run_layers(model, h: torch.Tensor, meta) -> torch.Tensor:
    B, T, _ = h.shape
    plans = model.__dict__.setdefault("_dstep_plans", {})
    key = (B, T, int(meta.seq_hint or 0))
    plan: Optional[DecodeStepPlan] = plans.get(key)
    if plan is None:
        plan = plans[key] = DecodeStepPlan(model, B, T, int(meta.seq_hint or 0), h.device)
    a0 = model.layers[0].self_attn
    cos, sin = a0._rope(meta)
    if meta.lines is None:
        meta.lines = model.kv_mgr.lines_for(meta.seq_ids)
    return plan.run(h, meta, cos.float() if cos.dtype != torch.float32 else cos, sin.float() if sin.dtype != torch.float32 else sin,
                    meta.lines.to(torch.int32))
