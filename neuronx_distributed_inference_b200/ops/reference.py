"""Plain-PyTorch definitions of every op in ``ops``.

Two jobs: (1) the CPU execution path (``on_cpu`` / gloo tests — the role ``to_cpu()`` plays in
the reference, application_base.py:556-628), (2) the fp32 oracle that each CUDA kernel's
numerics test compares against.  Shapes: B batch, T active tokens, S cache length,
Hq/Hkv heads, D head dim.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def rmsnorm(x, weight, eps: float, offset: float = 0.0, residual=None):
    """y = x * rsqrt(mean(x^2)+eps) * (offset + weight), statistics in fp32
    (reference custom_calls.py:8-36; Gemma uses offset=1)."""
    if residual is not None:
        x = x + residual
        residual = x
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    y = xf * torch.rsqrt(var + eps)
    if weight is not None:
        y = y * (weight.float() + offset)
    y = y.to(x.dtype)
    return y if residual is None else (y, residual)


def _dequant_weight(w, scale):
    if scale is None:
        return w
    wf = w.float()
    s = scale.float()
    if s.dim() == 1:
        s = s.unsqueeze(-1)
    elif s.dim() == 2 and s.shape != wf.shape and s.shape[-1] != 1:
        # blockwise [out/bs0, in/bs1] -> expand
        r0, r1 = wf.shape[0] // s.shape[0], wf.shape[1] // s.shape[1]
        s = s.repeat_interleave(r0, 0).repeat_interleave(r1, 1)
    return wf * s


def linear(x, w, bias=None, norm_weight=None, norm_eps: float = 1e-6, norm_offset: float = 0.0,
           act: Optional[str] = None, scale=None):
    """y = act(norm(x) @ dequant(w)^T + bias)."""
    if norm_weight is not None:
        x = rmsnorm(x, norm_weight, norm_eps, norm_offset)
    if scale is not None:
        w = _dequant_weight(w, scale).to(x.dtype)
    y = F.linear(x, w, bias)
    if act is not None:
        y = activation(y, act)
    return y


def activation(y, act: str):
    if act == "silu_mul":
        g, u = y.chunk(2, -1)
        return F.silu(g.float()).to(y.dtype) * u
    if act == "gelu_mul":
        g, u = y.chunk(2, -1)
        return F.gelu(g.float()).to(y.dtype) * u
    if act == "gelu_tanh_mul":
        g, u = y.chunk(2, -1)
        return F.gelu(g.float(), approximate="tanh").to(y.dtype) * u
    if act == "silu":
        return F.silu(y)
    if act == "gelu":
        return F.gelu(y)
    if act == "gelu_tanh":
        return F.gelu(y, approximate="tanh")
    if act == "relu":
        return F.relu(y)
    if act == "relu2":
        return F.relu(y).square()
    raise ValueError(act)


def rotate_half(x):
    x1, x2 = x.chunk(2, -1)
    return torch.cat((-x2, x1), -1)


def apply_rope(x, cos, sin, interleaved: bool = False):
    """x [B,T,H,D]; cos/sin [B,T,R/2] (R = rotary dim <= D).  Half-rotation (HF Llama,
    reference attention/utils.py:233-249) or interleaved pairs (Llama-4 / GPT-J)."""
    R = cos.shape[-1] * 2
    xr, xp = x[..., :R], x[..., R:]
    c = cos.unsqueeze(2).float()
    s = sin.unsqueeze(2).float()
    xf = xr.float()
    if interleaved:
        x1, x2 = xf[..., 0::2], xf[..., 1::2]
        o = torch.stack((x1 * c - x2 * s, x2 * c + x1 * s), -1).flatten(-2)
    else:
        x1, x2 = xf[..., : R // 2], xf[..., R // 2:]
        o = torch.cat((x1 * c - x2 * s, x2 * c + x1 * s), -1)
    o = o.to(x.dtype)
    return o if xp.shape[-1] == 0 else torch.cat((o, xp), -1)


def _expand_kv(k, n_rep):
    if n_rep == 1:
        return k
    B, H, S, D = k.shape
    return k[:, :, None].expand(B, H, n_rep, S, D).reshape(B, H * n_rep, S, D)


def attention_with_mask(q, k, v, mask, scale, sinks=None, softcap: Optional[float] = None):
    """q [B,Hq,T,D], k/v [B,Hkv,S,D], boolean mask [B,1|Hq,T,S] (True = attend).  fp32 softmax.
    ``sinks`` [Hq]: learned per-head logit joining the softmax denominator (GPT-OSS,
    reference attention_base.py learned-sink variant)."""
    n_rep = q.shape[1] // k.shape[1]
    k = _expand_kv(k, n_rep)
    v = _expand_kv(v, n_rep)
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if softcap:
        s = torch.tanh(s / softcap) * softcap
    s = s.masked_fill(~mask, float("-inf"))
    if sinks is not None:
        sk = sinks.float().view(1, -1, 1, 1).expand(s.shape[0], -1, s.shape[2], 1)
        p = torch.softmax(torch.cat((s, sk), -1), -1)[..., :-1]
    else:
        m = s.amax(-1, keepdim=True)
        m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
        e = torch.exp(s - m)
        p = e / e.sum(-1, keepdim=True).clamp_min(1e-30)
    return torch.matmul(p, v.float()).to(q.dtype)


def build_mask(q_pos, kv_len: int, window: Optional[int] = None, chunk: Optional[int] = None,
               key_valid=None, device=None):
    """Causal / sliding-window / chunked mask.  q_pos [B,T] absolute positions; key j visible iff
    j <= pos (causal), j > pos - window (SWA, reference model_base.py:187-376 mask builders),
    floor(j/chunk) == floor(pos/chunk) (Llama-4 chunked attention).  -> [B,1,T,S] bool."""
    j = torch.arange(kv_len, device=q_pos.device).view(1, 1, -1)
    p = q_pos.unsqueeze(-1)
    m = j <= p
    if window:
        m = m & (j > p - window)
    if chunk:
        m = m & ((j // chunk) == (p // chunk))
    if key_valid is not None:
        m = m & key_valid.bool().unsqueeze(1)
    return m.unsqueeze(1)


def attention_prefill(q, k, v, scale, causal: bool = True, window=None, chunk=None, key_valid=None,
                      q_pos=None, sinks=None, softcap=None):
    """q [B,T,Hq,D], k/v [B,S,Hkv,D] (new tokens, T==S unless prefix) -> [B,T,Hq,D]."""
    B, T = q.shape[:2]
    S = k.shape[1]
    if q_pos is None:
        q_pos = torch.arange(S - T, S, device=q.device).unsqueeze(0).expand(B, T)
    if causal:
        mask = build_mask(q_pos, S, window, chunk, key_valid)
    else:
        mask = torch.ones(B, 1, T, S, dtype=torch.bool, device=q.device)
        if key_valid is not None:
            mask = mask & key_valid.bool().view(B, 1, 1, S)
    o = attention_with_mask(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), mask, scale, sinks, softcap)
    return o.transpose(1, 2)


def kv_append(k_cache, v_cache, k_new, v_new, seq_ids, positions):
    """Write k_new/v_new [B,T,Hkv,D] into caches [L,Hkv,S,D] at (line=seq_ids[b], pos=positions[b,t]).
    Out-of-range lines/positions are skipped (reference K10 OOB-skip semantics)."""
    B, T = positions.shape
    L, H, S, D = k_cache.shape
    line = seq_ids.view(B, 1).expand(B, T).reshape(-1).long()
    pos = positions.reshape(-1).long()
    ok = (line >= 0) & (line < L) & (pos >= 0) & (pos < S)
    line, pos = line[ok], pos[ok]
    kn = k_new.reshape(B * T, H, D)[ok].to(k_cache.dtype)
    vn = v_new.reshape(B * T, H, v_new.shape[-1])[ok].to(v_cache.dtype)   # V may have its own head dim (MLA)
    k_cache[line, :, pos] = kn
    v_cache[line, :, pos] = vn


def attention_decode(q, k_cache, v_cache, seq_ids, positions, scale, window=None, chunk=None,
                     sinks=None, active_mask=None, softcap=None, k_scale=None, v_scale=None, active_base=None):
    """Attend q [B,T,Hq,D] over cache lines (already containing the active tokens).
    Token t of row b sees cache slots j <= positions[b,t] (+window/chunk).  ``active_mask`` [B,T,M] (token-tree
    speculation, M >= T) gives the visibility of the M slots starting at ``active_base`` [B,1] (default: the slot of the
    first active token); slots before the base are fully visible, slots after ``base+M`` are hidden."""
    B, T = positions.shape
    S = k_cache.shape[2]
    lines = seq_ids.long().clamp(0, k_cache.shape[0] - 1)
    k = k_cache[lines]
    v = v_cache[lines]
    if k.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        k = k.to(q.dtype) * (1.0 if k_scale is None else k_scale)
        v = v.to(q.dtype) * (1.0 if v_scale is None else v_scale)
    mask = build_mask(positions, S, window, chunk)
    if active_mask is not None:
        M = active_mask.shape[-1]
        base = (positions[:, :1] if active_base is None else active_base.view(B, 1)).long().unsqueeze(-1)   # [B,1,1]
        j = torch.arange(S, device=q.device).view(1, 1, S)
        prior = (j < base).expand(B, T, S)
        idx = (j - base).clamp(0, M - 1).expand(B, T, S)
        act = torch.gather(active_mask.bool().expand(B, T, M), 2, idx) & (j >= base) & (j < base + M)
        mask = (prior | act).unsqueeze(1)
        if window:
            mask = mask & build_mask(positions, S, window, None)
    o = attention_with_mask(q.transpose(1, 2), k, v, mask, scale, sinks, softcap)
    return o.transpose(1, 2)


def paged_kv_append(k_cache, v_cache, k_new, v_new, slot_mapping):
    """caches [num_blocks, block_size, Hkv, D]; slot = block*block_size+offset, -1 = skip
    (reference block_kv_cache_manager.py:268-328)."""
    nb, bs, H, D = k_cache.shape
    slots = slot_mapping.reshape(-1).long()
    ok = slots >= 0
    kc = k_cache.view(nb * bs, H, D)
    vc = v_cache.view(nb * bs, H, D)
    kc[slots[ok]] = k_new.reshape(-1, H, D)[ok].to(k_cache.dtype)
    vc[slots[ok]] = v_new.reshape(-1, H, D)[ok].to(v_cache.dtype)


def paged_attention_decode(q, k_cache, v_cache, block_table, positions, scale, window=None, sinks=None):
    """q [B,T,Hq,D]; block_table [B, max_blocks]; positions [B,T]."""
    nb, bs, H, D = k_cache.shape
    B = q.shape[0]
    bt = block_table.long().clamp_min(0)
    k = k_cache[bt].reshape(B, -1, H, D).transpose(1, 2)
    v = v_cache[bt].reshape(B, -1, H, D).transpose(1, 2)
    mask = build_mask(positions, k.shape[2], window)
    o = attention_with_mask(q.transpose(1, 2), k, v, mask, scale, sinks)
    return o.transpose(1, 2)


def sample(logits, top_k, top_p, temperature, rand=None, global_topk: int = 256):
    """On-device sampling semantics of the reference (sampling.py:241-464): restrict to the global
    top-k (k<=global_topk), per-row top-k mask, temperature, top-p mask on the cumulative softmax,
    softmax, inverse-CDF draw with ``rand`` [B] uniform (0.5 when deterministic).
    logits [B,V] fp32; top_k [B] int (<=0 => global_topk); top_p, temperature [B]. -> tokens [B]."""
    B, V = logits.shape
    K = min(global_topk, V)
    vals, idx = torch.topk(logits.float(), K, dim=-1)
    k = torch.where(top_k <= 0, torch.full_like(top_k, K), top_k).clamp(max=K).view(B, 1)
    ar = torch.arange(K, device=logits.device).view(1, K)
    vals = vals.masked_fill(ar >= k, float("-inf"))
    t = temperature.float().view(B, 1)
    greedy = (t == 0) | (k == 1)
    vals = vals / torch.where(t == 0, torch.ones_like(t), t)
    probs = torch.softmax(vals, -1)
    cum = torch.cumsum(probs, -1)
    keep = (cum - probs) < top_p.float().view(B, 1)  # always keeps the first token
    probs = torch.where(keep, probs, torch.zeros_like(probs))
    probs = probs / probs.sum(-1, keepdim=True)
    cdf = torch.cumsum(probs, -1)
    if rand is None:
        rand = torch.full((B,), 0.5, device=logits.device)
    r = rand.float().view(B, 1) * cdf[:, -1:]
    choice = (cdf < r).sum(-1).clamp(max=K - 1)
    choice = torch.where(greedy.view(B), torch.zeros_like(choice), choice)
    return idx.gather(1, choice.view(B, 1)).view(B)


def argmax(logits):
    return logits.float().argmax(-1)


def moe_route(router_logits, top_k: int, act: str = "softmax", normalize: bool = True,
              act_over_topk: bool = False):
    """-> (weights [N,top_k] fp32, expert ids [N,top_k]).  softmax-then-topk (Mixtral/DBRX/Qwen3),
    sigmoid (Llama-4) or topk-then-softmax (GPT-OSS ``apply_act_fn_over_topk``)."""
    x = router_logits.float()
    if act_over_topk:
        v, i = torch.topk(x, top_k, -1)
        w = torch.softmax(v, -1) if act == "softmax" else torch.sigmoid(v)
        return w, i
    p = torch.softmax(x, -1) if act == "softmax" else torch.sigmoid(x)
    w, i = torch.topk(p, top_k, -1)
    if normalize and act == "softmax":
        w = w / w.sum(-1, keepdim=True)
    return w, i


def moe_experts(x, w_gate_up, w_down, topk_w, topk_i, act: str = "silu_mul", expert_offset: int = 0,
                gate_up_bias=None, down_bias=None, act_fn=None, scale_input: bool = False, gate_up_scale=None, down_scale=None):
    """Dropless expert MLPs.  x [N,H]; w_gate_up [E_local, 2I, H] ([gate; up] rows, K-major like nn.Linear — the
    reference stores [E,H,2I]; K-major is what TMA/UMMA and the GEMV kernels stream); w_down
    [E_local, H, I]; experts owned here are [expert_offset, expert_offset+E_local).
    -> partial output [N,H] (sum over owned experts; all-reduce across expert/tensor shards outside)."""
    N, H = x.shape
    E = w_gate_up.shape[0]
    out = torch.zeros(N, H, dtype=torch.float32, device=x.device)
    for e in range(E):
        sel = (topk_i == e + expert_offset)
        tok = sel.any(-1).nonzero().flatten()
        if tok.numel() == 0:
            continue
        wt = (topk_w * sel).sum(-1)[tok]
        xin = x[tok]
        if scale_input:   # Llama-4 "early expert affinity modulation": the affinity scales the expert INPUT
            xin = (xin.float() * wt.unsqueeze(-1)).to(x.dtype)
            wt = torch.ones_like(wt)
        h = F.linear(xin, w_gate_up[e].to(x.dtype))
        if gate_up_scale is not None:      # weight-only 8-bit experts: per-(expert, output channel) scale applied to the GEMM output
            h = (h.float() * gate_up_scale[e].float()).to(x.dtype)
        if gate_up_bias is not None:
            h = h + gate_up_bias[e].to(h.dtype)
        h = act_fn(h) if act_fn is not None else activation(h, act)
        y = F.linear(h, w_down[e].to(x.dtype))
        if down_scale is not None:
            y = (y.float() * down_scale[e].float()).to(x.dtype)
        if down_bias is not None:
            y = y + down_bias[e].to(y.dtype)
        out[tok] += y.float() * wt.unsqueeze(-1)
    return out.to(x.dtype)


def quantize_per_channel(w, dtype=torch.int8, axis: int = 0):
    """Symmetric per-output-channel quantisation.  -> (q, scale fp32 [out])."""
    wf = w.float()
    amax = wf.abs().amax(dim=1 - axis if w.dim() == 2 else -1, keepdim=True).clamp_min(1e-12)
    qmax = 127.0 if dtype == torch.int8 else (448.0 if dtype == torch.float8_e4m3fn else 57344.0)
    scale = amax / qmax
    q = wf / scale
    q = q.round().clamp(-qmax, qmax).to(torch.int8) if dtype == torch.int8 else q.clamp(-qmax, qmax).to(dtype)
    return q, scale.squeeze(-1)


def quantize_per_tensor(w, dtype=torch.int8):
    wf = w.float()
    qmax = 127.0 if dtype == torch.int8 else (448.0 if dtype == torch.float8_e4m3fn else 57344.0)
    scale = wf.abs().max().clamp_min(1e-12) / qmax
    q = wf / scale
    q = q.round().clamp(-qmax, qmax).to(torch.int8) if dtype == torch.int8 else q.clamp(-qmax, qmax).to(dtype)
    return q, scale.reshape(1)


def quantize_blockwise(w, block=(128, 128), dtype=torch.float8_e4m3fn):
    out, inp = w.shape
    b0, b1 = block
    assert out % b0 == 0 and inp % b1 == 0
    wf = w.float().view(out // b0, b0, inp // b1, b1)
    qmax = 448.0 if dtype == torch.float8_e4m3fn else 127.0
    scale = wf.abs().amax((1, 3), keepdim=True).clamp_min(1e-12) / qmax
    q = (wf / scale)
    q = q.round().clamp(-qmax, qmax).to(torch.int8) if dtype == torch.int8 else q.clamp(-qmax, qmax).to(dtype)
    return q.view(out, inp), scale.view(out // b0, inp // b1)


def rmsnorm_quant(x, weight, eps, clamp: float = float("inf")):
    """RMSNorm + dynamic per-row fp8(e4m3) quantisation (reference kernel K6)."""
    y = rmsnorm(x, weight, eps).float()
    if math.isfinite(clamp):
        y = y.clamp(-clamp, clamp)
    amax = y.abs().amax(-1, keepdim=True).clamp_min(1e-12)
    scale = amax / 448.0
    return (y / scale).to(torch.float8_e4m3fn), scale
