"""Op dispatch: hand-written sm_100a kernels on CUDA tensors, PyTorch reference on CPU.

There is exactly one accelerated backend (the in-tree ``_C`` extension built from ``csrc/``);
there is no multi-backend layer.  A CUDA tensor with the extension missing is an error (set
``NXDI_B200_ALLOW_TORCH_FALLBACK=1`` to run the PyTorch definitions on GPU for debugging).
``stats`` counts kernel launches issued through this module (bench.py reports them).
"""
from __future__ import annotations

import os
from collections import Counter
from typing import Optional

import torch

from . import reference as ref
from ._ext import load_extension, extension_available

stats = Counter()
_ALLOW_FALLBACK = os.environ.get("NXDI_B200_ALLOW_TORCH_FALLBACK", "0") == "1"
GEMV_MAX_TOKENS = 8      # CUDA-core weight-streaming kernel up to here
_KERNELS_ENABLED = True
_TCGEN05_GEMM = os.environ.get("NXDI_B200_TCGEN05_GEMM", "1") == "1"   # 0: cuBLAS for T > 8 (A/B baseline)


def set_kernels_enabled(flag: bool):
    """Debug switch: run the PyTorch definitions on CUDA tensors (accuracy triage)."""
    global _KERNELS_ENABLED
    _KERNELS_ENABLED = flag


def _C():
    return load_extension()


def _use_cuda(x: torch.Tensor) -> bool:
    if not x.is_cuda or not _KERNELS_ENABLED:
        return False
    if extension_available():
        return True
    if _ALLOW_FALLBACK:
        return False
    raise RuntimeError("CUDA tensor but the sm_100a extension is not built: run "
                       "`python -c 'import __graft_entry__ as g; g.build()'` "
                       "(or set NXDI_B200_ALLOW_TORCH_FALLBACK=1)")


_FAST_DTYPES = (torch.bfloat16,)


# --------------------------------------------------------------------------------------------
def rmsnorm(x, weight, eps: float, offset: float = 0.0, residual=None):
    if _use_cuda(x) and x.dtype in _FAST_DTYPES and x.shape[-1] % 8 == 0 and weight is not None \
            and weight.dtype == x.dtype:
        stats["rmsnorm"] += 1
        x2 = x.reshape(-1, x.shape[-1])
        if residual is not None:
            r2 = residual.reshape(-1, x.shape[-1])
            y, r = _C().rmsnorm(x2, weight, eps, offset, r2)
            return y.view(x.shape), r.view(x.shape)
        y, _ = _C().rmsnorm(x2, weight, eps, offset, None)
        return y.view(x.shape)
    return ref.rmsnorm(x, weight, eps, offset, residual)


_ACT_CODES = {None: 0, "silu_mul": 1, "gelu_tanh_mul": 2, "gelu_mul": 3}


def staging_for(group, x, w):
    """Output tensor for a row-parallel GEMM inside the group's symmetric staging area (partial sums that an in-switch collective
    consumes next), or None when the group has no heap / the shape does not qualify."""
    heap = getattr(group, "heap", None)
    if heap is None or not x.is_cuda or x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        return None
    shape = tuple(x.shape[:-1]) + (w.shape[0],)
    T = x.numel() // x.shape[-1]
    if T <= GEMV_MAX_TOKENS or not heap.usable(torch.empty(0, dtype=torch.bfloat16, device=x.device), None, shape):
        return None
    return heap.staging(shape)


def linear(x, w, bias=None, norm_weight=None, norm_eps: float = 1e-6, norm_offset: float = 0.0,
           act: Optional[str] = None, scale=None, residual=None, out=None):
    """y = act(rmsnorm(x) @ w^T + bias) (+ residual).  w: [N,K] bf16, or int8/fp8 with per-channel ``scale``.
    ``out``: preallocated result (tcgen05 GEMM path only; used to land partial sums in the symmetric heap)."""
    if _use_cuda(x) and x.dtype in _FAST_DTYPES:
        K = x.shape[-1]
        T = x.numel() // K
        N = w.shape[0]
        wq = w.dtype in (torch.int8, torch.float8_e4m3fn)
        ok_w = (w.dtype == x.dtype and scale is None) or (wq and scale is not None and scale.dim() == 1)
        if ok_w and act in _ACT_CODES and w.is_contiguous():
            x2 = x.reshape(T, K)
            if wq and T <= GEMV_MAX_TOKENS and K % 16 == 0 and w.dim() == 2:
                stats["qgemv"] += 1
                r2 = residual.reshape(T, -1).contiguous() if (residual is not None and act is None) else None
                y = _C().gemv(x2, w, bias, norm_weight, norm_eps, norm_offset, _ACT_CODES[act], scale.float().contiguous(), r2)
                y = y.view(*x.shape[:-1], y.shape[-1])
                return y if (residual is None or r2 is not None) else y + residual
            if not wq and T <= GEMV_MAX_TOKENS and K % 64 == 0 and (K % 256 == 0 or _C().gemv2_supported(T, K)):
                stats["gemv"] += 1
                r2 = residual.reshape(T, -1) if (residual is not None and act is None) else None
                y = _C().gemv(x2, w, bias, norm_weight, norm_eps, norm_offset, _ACT_CODES[act], scale, r2)
                y = y.view(*x.shape[:-1], y.shape[-1])
                return y if (residual is None or r2 is not None) else y + residual
            n_out = N // 2 if act is not None else N
            if (wq and _ACT_QUANT and w.dtype == torch.float8_e4m3fn and T > GEMV_MAX_TOKENS and K % 128 == 0 and n_out % 8 == 0
                    and w.dim() == 2 and K <= 16384):
                # W8A8: RMSNorm + per-token fp8 quantisation in one kernel, then the fp8 tensor-core GEMM (2x the bf16 rate);
                # acc * a_scale[row] * w_scale[col] in the epilogue
                stats["gemm_fp8"] += 1
                xq, a_s = _C().rmsnorm_quant(x2.contiguous(), norm_weight, float(norm_eps), float(norm_offset), float("inf"))
                r2 = residual.reshape(T, -1).contiguous() if (residual is not None and act is None) else None
                y = _C().gemm_fp8(xq, a_s, w, scale.float().contiguous(), bias, _ACT_CODES[act], r2)
                y = y.view(*x.shape[:-1], y.shape[-1])
                return y if (residual is None or r2 is not None) else y + residual
            if (wq and T > GEMV_MAX_TOKENS and w.dim() == 2 and K % 64 == 0 and n_out % 8 == 0 and _TCGEN05_GEMM
                    and x.dtype == torch.bfloat16):
                # weight-only 8-bit prefill: one expansion pass into a shared bf16 scratch (stays in L2), then the bf16 tensor-core GEMM
                stats["dequant_gemm"] += 1
                w, scale, wq = _C().dequant_bf16(w, scale.float().contiguous(), _dequant_scratch(w)), None, False
            if not wq and n_out % 8 == 0 and K % 64 == 0 and _TCGEN05_GEMM and (T > GEMV_MAX_TOKENS or K % 64 != 0):
                if norm_weight is not None:
                    x2 = rmsnorm(x2, norm_weight, norm_eps, norm_offset)
                stats["gemm_tcgen05"] += 1
                r2 = residual.reshape(T, -1) if (residual is not None and act is None) else None
                o2 = out.view(T, -1) if out is not None else None
                y = _C().gemm(x2.contiguous(), w, bias, _ACT_CODES[act], r2, o2)
                y = out if out is not None else y.view(*x.shape[:-1], y.shape[-1])
                return y if (residual is None or r2 is not None) else y + residual
    y = ref.linear(x, w, bias, norm_weight, norm_eps, norm_offset, act, scale)
    y = y if residual is None else y + residual
    if out is not None:
        out.copy_(y)
        return out
    return y


_DEQ_SCRATCH = {}
_DEQ_RETIRED = []


def _dequant_scratch(w):
    """One bf16 scratch per device, grown to the largest 8-bit weight seen (consumed by the GEMM enqueued right after the expansion;
    stream order makes the reuse by the next layer safe)."""
    buf = _DEQ_SCRATCH.get(w.device)
    if buf is None or buf.numel() < w.numel():
        if buf is not None:
            _DEQ_RETIRED.append(buf)      # never freed: CUDA graphs captured with the smaller scratch still hold its address
        buf = torch.empty(w.numel(), dtype=torch.bfloat16, device=w.device)
        _DEQ_SCRATCH[w.device] = buf
    return buf


def linear_allreduce(x, w, bias, group, residual=None, reduce_dtype=None, scale=None):
    """Row-parallel GEMM -> all-reduce (+bias, +residual).  T <= 8 tokens on CUDA with a
    symmetric workspace attached to the group: ONE kernel (GEMV, P2P stores of partials into every
    peer, flag wait, reduce, residual add).  Otherwise GEMM then NCCL/gloo all-reduce."""
    from ..parallel import mappings
    K = x.shape[-1]
    T = x.numel() // K
    wq = w.dtype in (torch.int8, torch.float8_e4m3fn) and scale is not None and scale.dim() == 1 and w.dim() == 2
    if (_use_cuda(x) and x.dtype in _FAST_DTYPES and group.symm is not None and T <= GEMV_MAX_TOKENS
            and w.shape[0] <= group.symm.n_max
            and K % 64 == 0 and w.is_contiguous() and ((w.dtype == x.dtype and scale is None) or wq)
            and _C().gemv2_supported(T, K, (1 if w.dtype == torch.int8 else 2) if wq else 0)
            and reduce_dtype in (None, torch.float32)):
        stats["gemv_allreduce"] += 1
        y = group.symm.gemv_allreduce(x.reshape(T, K), w, bias, residual.reshape(T, -1) if residual is not None else None,
                                      scale.float().contiguous() if wq else None)
        return y.view(*x.shape[:-1], y.shape[-1])
    heap = getattr(group, "heap", None)
    if heap is not None and scale is None and reduce_dtype in (None, torch.float32, x.dtype) and T > GEMV_MAX_TOKENS \
            and heap.gemm_ar_usable(x, w):
        # prefill: tcgen05 GEMM -> per-tile hand-off -> in-switch reduce of the owned tiles -> multicast of the result, ONE kernel
        stats["gemm_all_reduce"] += 1
        y = heap.gemm_all_reduce(x.reshape(T, K), w, bias if group.rank == 0 else None, residual)
        return y.view(*x.shape[:-1], w.shape[0])
    if heap is not None and scale is None:
        # prefill: partial sums land in the symmetric staging area (bias on rank 0 only), one in-switch all-reduce adds the residual
        y = linear(x, w, bias if group.rank == 0 else None, out=staging_for(group, x, w))
        return mappings.all_reduce(y, group, reduce_dtype=reduce_dtype, residual=residual)
    y = linear(x, w, None, scale=scale)
    y = mappings.all_reduce(y, group, reduce_dtype=reduce_dtype)
    if bias is not None:
        y = y + bias
    if residual is not None:
        y = y + residual
    return y


def apply_rope(x, cos, sin, interleaved: bool = False):
    return ref.apply_rope(x, cos, sin, interleaved)


def rope_kv_append(qkv, cos, sin, k_cache, v_cache, seq_ids, positions, n_q, n_kv, head_dim,
                   interleaved: bool = False, q_norm=None, k_norm=None, norm_eps: float = 1e-6):
    """Split fused qkv [B,T,(n_q+2n_kv)*D], optional per-head q/k RMSNorm (Qwen3), RoPE on q,k,
    write k,v into the contiguous cache at (seq_ids[b], positions[b,t]).  Returns q [B,T,n_q,D].
    ONE kernel on CUDA (csrc/rope_kv.cu)."""
    B, T = positions.shape
    D = head_dim
    if (_use_cuda(qkv) and qkv.dtype in _FAST_DTYPES and k_cache.dtype == qkv.dtype and D in (64, 128, 256)
            and cos.shape[-1] * 2 == D and not interleaved):
        stats["rope_kv_append"] += 1
        return _C().rope_kv_append(qkv.reshape(B, T, -1), cos.contiguous(), sin.contiguous(), k_cache, v_cache,
                                   seq_ids.to(torch.int32), positions.to(torch.int32), n_q, n_kv, D,
                                   q_norm, k_norm, norm_eps)
    q, k, v = qkv.reshape(B, T, n_q + 2 * n_kv, D).split([n_q, n_kv, n_kv], dim=2)
    if q_norm is not None:
        q = ref.rmsnorm(q, q_norm, norm_eps)
        k = ref.rmsnorm(k, k_norm, norm_eps)
    q = ref.apply_rope(q, cos, sin, interleaved)
    k = ref.apply_rope(k, cos, sin, interleaved)
    kv_append(k_cache, v_cache, k, v, seq_ids, positions)
    return q


def rope_kv_split_append(qkv, cos, sin, k_cache, v_cache, seq_ids, positions, n_q, n_kv, head_dim, q_norm=None, k_norm=None,
                         norm_eps: float = 1e-6):
    """Prefill: split the fused QKV projection, per-head q/k RMSNorm, RoPE, cache write AND the contiguous rotated k / v of the new
    tokens for the flash kernel — ONE kernel (csrc/rope_kv.cu) instead of the split / rotate / cat / append chain.
    -> (q [B,T,n_q,D], k [B,T,n_kv,D], v [B,T,n_kv,D]) or None when the kernel does not apply."""
    B, T = positions.shape
    D = head_dim
    if (_use_cuda(qkv) and qkv.dtype == torch.bfloat16 and k_cache.dtype == qkv.dtype and D in (64, 128, 256) and cos is not None
            and cos.shape[-1] * 2 == D):
        stats["rope_kv_split_append"] += 1
        return _C().rope_kv_split_append(qkv.reshape(B, T, -1).contiguous(), cos.contiguous(), sin.contiguous(), k_cache, v_cache,
                                         seq_ids.to(torch.int32), positions.to(torch.int32).contiguous(), n_q, n_kv, D, q_norm, k_norm,
                                         norm_eps)
    return None


def kv_append(k_cache, v_cache, k_new, v_new, seq_ids, positions):
    if (_use_cuda(k_new) and k_new.dtype in _FAST_DTYPES and k_cache.dtype == k_new.dtype
            and (k_new.shape[-1] * k_new.element_size()) % 16 == 0 and v_new.shape[-1] == k_new.shape[-1]):
        stats["kv_append"] += 1
        _C().kv_append(k_cache, v_cache, k_new.contiguous(), v_new.contiguous(), seq_ids.to(torch.int32),
                       positions.to(torch.int32))
        return
    ref.kv_append(k_cache, v_cache, k_new, v_new, seq_ids, positions)


def attention_decode(q, k_cache, v_cache, seq_ids, positions, scale, window=None, chunk=None, sinks=None,
                     active_mask=None, softcap=None, k_scale=None, v_scale=None, seq_hint: int = 0, active_base=None):
    """``seq_hint``: upper bound on the live context (the TKG bucket) used only to size the split-KV grid."""
    D = q.shape[-1]
    if (_use_cuda(q) and q.dtype in _FAST_DTYPES and k_cache.dtype == q.dtype and D in (64, 128)
            and active_mask is None and chunk is None and softcap is None
            and q.shape[1] * (q.shape[2] // k_cache.shape[1]) <= 64):
        stats["attn_decode"] += 1
        return _C().attention_decode(q.contiguous(), k_cache, v_cache, seq_ids.to(torch.int32),
                                     positions.to(torch.int32).contiguous(), float(scale), int(window or 0), sinks,
                                     int(seq_hint))
    return ref.attention_decode(q, k_cache, v_cache, seq_ids, positions, scale, window, chunk, sinks,
                                active_mask, softcap, k_scale, v_scale, active_base)


def rope_attention_decode(qkv, cos, sin, k_cache, v_cache, seq_ids, write_positions, positions, n_q, n_kv, head_dim, scale,
                          window=None, sinks=None, q_norm=None, k_norm=None, norm_eps: float = 1e-6, seq_hint: int = 0):
    """Decode attention straight from the QKV projection: per-head q/k RMSNorm, RoPE, cache append and split-KV flash
    decode in ONE kernel (the append is the attention kernel's prologue).  qkv [B,T,(n_q+2n_kv)D] -> [B,T,n_q,D].
    Falls back to ``rope_kv_append`` + ``attention_decode`` where the fused kernel does not apply."""
    B, T = positions.shape
    D = head_dim
    if (_use_cuda(qkv) and qkv.dtype in _FAST_DTYPES and k_cache.dtype == qkv.dtype and D in (64, 128) and cos.shape[-1] * 2 == D
            and T * (n_q // n_kv) <= 64 and os.environ.get("NXDI_B200_FUSED_ROPE_ATTN", "1") != "0"):
        stats["rope_attn_decode"] += 1
        return _C().rope_attention_decode(qkv.reshape(B, T, -1), cos.contiguous(), sin.contiguous(), k_cache, v_cache,
                                          seq_ids.to(torch.int32), write_positions.to(torch.int32).contiguous(),
                                          positions.to(torch.int32).contiguous(), n_q, n_kv, D, float(scale), int(window or 0),
                                          sinks, q_norm, k_norm, norm_eps, int(seq_hint))
    q = rope_kv_append(qkv, cos, sin, k_cache, v_cache, seq_ids, write_positions, n_q, n_kv, D, False, q_norm, k_norm, norm_eps)
    return attention_decode(q, k_cache, v_cache, seq_ids, positions, scale, window, None, sinks, seq_hint=seq_hint)


_ATTN_TC = os.environ.get("NXDI_B200_ATTN_TC", "1") != "0"   # 0: the mma.sync flash kernel for head_dim 128 too (A/B baseline)


def attention_prefill(q, k, v, scale, causal: bool = True, window=None, chunk=None, key_valid=None,
                      q_pos=None, sinks=None, softcap=None):
    D = q.shape[-1]
    if (_use_cuda(q) and q.dtype in _FAST_DTYPES and D == 128 and _ATTN_TC and chunk is None and (causal or not window)
            and key_valid is None and q_pos is None and q.shape[1] == k.shape[1] and k.dtype == q.dtype):
        # Blackwell path: tcgen05 QK^T / PV with TMEM accumulators, TMA tiles (csrc/attention_tc.cu); soft-cap is an argument
        stats["attn_prefill_tc"] += 1
        return _C().attention_prefill_tc(q.contiguous(), k.contiguous(), v.contiguous(), float(scale), int(window or 0), sinks,
                                         bool(causal), float(softcap or 0.0))
    if (_use_cuda(q) and q.dtype in _FAST_DTYPES and D in (64, 128) and chunk is None and (causal or not window)
            and key_valid is None and q_pos is None and softcap is None and q.shape[1] == k.shape[1]):
        stats["attn_prefill"] += 1
        return _C().attention_prefill(q.contiguous(), k.contiguous(), v.contiguous(), float(scale),
                                      int(window or 0), sinks, bool(causal))
    return ref.attention_prefill(q, k, v, scale, causal, window, chunk, key_valid, q_pos, sinks, softcap)


def paged_kv_append(k_cache, v_cache, k_new, v_new, slot_mapping):
    if _use_cuda(k_new) and k_new.dtype in _FAST_DTYPES and k_cache.dtype == k_new.dtype:
        stats["paged_kv_append"] += 1
        _C().paged_kv_append(k_cache, v_cache, k_new.contiguous(), v_new.contiguous(),
                             slot_mapping.to(torch.int32).contiguous())
        return
    ref.paged_kv_append(k_cache, v_cache, k_new, v_new, slot_mapping)


def paged_attention_decode(q, k_cache, v_cache, block_table, positions, scale, window=None, sinks=None):
    D = q.shape[-1]
    if (_use_cuda(q) and q.dtype in _FAST_DTYPES and k_cache.dtype == q.dtype and D in (64, 128)
            and q.shape[1] * (q.shape[2] // k_cache.shape[2]) <= 64):
        stats["paged_attn_decode"] += 1
        return _C().paged_attention_decode(q.contiguous(), k_cache, v_cache, block_table.to(torch.int32).contiguous(),
                                           positions.to(torch.int32), float(scale), int(window or 0), sinks)
    return ref.paged_attention_decode(q, k_cache, v_cache, block_table, positions, scale, window, sinks)


def argmax(logits):
    if _use_cuda(logits) and logits.dim() == 2 and logits.dtype in (torch.float32, torch.bfloat16):
        stats["argmax"] += 1
        return _C().argmax(logits.contiguous(), [], None, 0, 0, 0, 0)
    return ref.argmax(logits)


def argmax_sharded(logits, group):
    """Arg-max over a vocabulary-sharded logits row block [B, V_local] -> GLOBAL token ids [B], ONE kernel: the per-rank
    winners are exchanged through the group's symmetric workspace inside the arg-max kernel (csrc/sampling.cu).  Returns None
    when the fused path does not apply (CPU, no workspace): the caller falls back to all-gather + glue."""
    symm = getattr(group, "symm", None)
    if (symm is None or not _use_cuda(logits) or logits.dim() != 2 or logits.dtype not in (torch.float32, torch.bfloat16)
            or logits.shape[0] > symm.ARGMAX_ROWS):
        return None
    stats["argmax_exchange"] += 1
    return symm.argmax(logits.contiguous())


def sample(logits, top_k, top_p, temperature, rand=None, global_topk: int = 256):
    if _use_cuda(logits) and logits.dim() == 2 and logits.dtype in (torch.float32, torch.bfloat16) \
            and global_topk <= 256:
        stats["topk_sample"] += 1
        B = logits.shape[0]
        if rand is None:
            rand = torch.full((B,), 0.5, device=logits.device)
        return _C().topk_sample(logits.contiguous(), top_k.to(torch.int32), top_p.float(), temperature.float(),
                                rand.float(), int(global_topk))
    return ref.sample(logits, top_k, top_p, temperature, rand, global_topk)


def moe_route(router_logits, top_k, act="softmax", normalize=True, act_over_topk=False):
    return ref.moe_route(router_logits, top_k, act, normalize, act_over_topk)


def moe_experts(x, w_gate_up, w_down, topk_w, topk_i, act="silu_mul", expert_offset=0,
                gate_up_bias=None, down_bias=None, act_fn=None, scale_input=False, gate_up_scale=None, down_scale=None):
    N = x.shape[0]
    if gate_up_scale is not None or down_scale is not None:
        # fp8 experts with dynamic activation quantisation asked for: W8A8 grouped GEMMs on the kind::f8f6f4 tensor-core path
        if (_use_cuda(x) and _MOE_GROUPED and _ACT_QUANT and N > GEMV_MAX_TOKENS and x.dtype == torch.bfloat16 and act in _MOE_ACTS
                and w_gate_up.dtype == torch.float8_e4m3fn and w_down.dtype == torch.float8_e4m3fn and act_fn is None
                and gate_up_scale is not None and down_scale is not None and gate_up_scale.dim() == 2 and down_scale.dim() == 2
                and x.shape[1] % 128 == 0 and w_down.shape[2] % 128 == 0 and max(x.shape[1], w_down.shape[2]) <= 16384
                and topk_i.dim() == 2 and topk_i.shape[1] <= 64 and w_gate_up.shape[0] <= 512
                and (gate_up_bias is None or gate_up_bias.dtype == x.dtype) and (down_bias is None or down_bias.dtype == x.dtype)):
            stats["moe_grouped_fp8"] += 1
            return _C().moe_grouped(x.contiguous(), w_gate_up.contiguous(), w_down.contiguous(), topk_w.float().contiguous(),
                                    topk_i.to(torch.int32).contiguous(), int(expert_offset), _MOE_ACTS[act], bool(scale_input),
                                    None if gate_up_bias is None else gate_up_bias.contiguous(),
                                    None if down_bias is None else down_bias.contiguous(),
                                    gate_up_scale.float().contiguous(), down_scale.float().contiguous())
        # weight-only 8-bit experts: dequantise the selected experts on the fly
        return ref.moe_experts(x, w_gate_up, w_down, topk_w, topk_i, act, expert_offset, gate_up_bias, down_bias, act_fn, scale_input,
                               gate_up_scale, down_scale)
    # activations with a kernel implementation: the named GLU ones, or a callable that names its kernel twin (``kernel_act``)
    kact = act if act_fn is None else getattr(act_fn, "kernel_act", None)
    if (_use_cuda(x) and x.dtype in _FAST_DTYPES and w_gate_up.dtype == x.dtype and act == "silu_mul"
            and act_fn is None and gate_up_bias is None and down_bias is None and N <= GEMV_MAX_TOKENS and not scale_input
            and w_gate_up.is_contiguous() and w_down.is_contiguous() and topk_i.dim() == 2
            and x.shape[1] % 8 == 0 and w_down.shape[2] % 8 == 0 and max(x.shape[1], w_down.shape[2]) * 2 <= 65536):
        stats["moe_decode"] += 1
        return _C().moe_decode(x.contiguous(), w_gate_up, w_down, topk_w.float().contiguous(),
                               topk_i.to(torch.int32).contiguous(), int(expert_offset))
    # prefill-sized batches: device-side permutation + grouped tcgen05 GEMMs (static shapes: replays under CUDA graphs)
    if (_use_cuda(x) and _MOE_GROUPED and x.dtype == torch.bfloat16 and w_gate_up.dtype == x.dtype and w_down.dtype == x.dtype
            and kact in _MOE_ACTS and topk_i.dim() == 2 and topk_i.shape[1] <= 64
            and w_gate_up.is_contiguous() and w_down.is_contiguous() and x.shape[1] % 64 == 0 and w_down.shape[2] % 64 == 0
            and w_gate_up.shape[0] <= 512
            and (gate_up_bias is None or gate_up_bias.dtype == x.dtype) and (down_bias is None or down_bias.dtype == x.dtype)):
        stats["moe_grouped"] += 1
        return _C().moe_grouped(x.contiguous(), w_gate_up, w_down, topk_w.float().contiguous(), topk_i.to(torch.int32).contiguous(),
                                int(expert_offset), _MOE_ACTS[kact], bool(scale_input),
                                None if gate_up_bias is None else gate_up_bias.contiguous(),
                                None if down_bias is None else down_bias.contiguous(), None, None)
    return ref.moe_experts(x, w_gate_up, w_down, topk_w, topk_i, act, expert_offset, gate_up_bias, down_bias,
                           act_fn, scale_input)


_MOE_GROUPED = os.environ.get("NXDI_B200_MOE_GROUPED", "1") == "1"
_MOE_ACTS = {"silu_mul": 1, "gelu_tanh_mul": 2, "gelu_mul": 3, "gpt_oss_glu": 4}


def rmsnorm_quant(x, weight, eps, clamp=float("inf"), offset: float = 0.0):
    """(fp8-e4m3 activations, per-row fp32 scale): RMSNorm (``weight`` may be None) fused with dynamic per-token quantisation.
    ONE kernel on CUDA (csrc/quant.cu); reference kernel K6 (modeling_llama.py:553-575)."""
    if _use_cuda(x) and x.dtype in _FAST_DTYPES and x.shape[-1] % 8 == 0 and x.shape[-1] <= 16384:
        stats["rmsnorm_quant"] += 1
        q, s = _C().rmsnorm_quant(x.reshape(-1, x.shape[-1]).contiguous(), weight, float(eps), float(offset), float(clamp))
        return q.view(x.shape), s.view(*x.shape[:-1], 1)
    return ref.rmsnorm_quant(x, weight, eps, clamp)


_ACT_QUANT = os.environ.get("NXDI_B200_FP8_ACT", "0") == "1"


def set_activation_quant(flag: bool):
    """W8A8: fp8 weights meet dynamically quantised fp8 activations on the tensor cores (tcgen05 kind::f8f6f4) for T > 8 tokens.
    Set from ``NeuronConfig.activation_quantization_type == "dynamic"`` / ``quantized_mlp_kernel_enabled``; off = weight-only."""
    global _ACT_QUANT
    _ACT_QUANT = bool(flag)


activation = ref.activation
build_mask = ref.build_mask
attention_with_mask = ref.attention_with_mask
quantize_per_channel = ref.quantize_per_channel
quantize_per_tensor = ref.quantize_per_tensor
quantize_blockwise = ref.quantize_blockwise
