"""Attention layer shared by every decoder model (the role of ``NeuronAttentionBase``,
reference modules/attention/attention_base.py:134-2488).

Data flow per layer (contiguous cache; the paged cache swaps the two cache ops):

    qkv   = Wqkv( rmsnorm(h) )                       one GEMV/GEMM, norm fused in the prologue
    q     = rope_kv_append(qkv, cos, sin, cache)     q/k norm + RoPE + cache write, one kernel
    o     = attention( q, cache[seq_ids], pos )      flash-decode (T<=16) or flash-prefill kernel
    h    += Wo(o)  [+ all-reduce]                    GEMV -> all-reduce -> +residual, one kernel

The reference computes decode attention as "prior (cache) + active (new tokens)" with a shared
softmax and writes the cache after the layer loop (attention_base.py:1410-1461); we append first
and attend over the cache, which is the same function of the same inputs and saves a pass.

Variants carried by flags rather than subclasses: sliding window, chunked attention (Llama-4),
learned sinks (GPT-OSS), q/k RMSNorm pre-RoPE (Qwen3) or L2 norm post-RoPE (Llama-4), clip_qkv
(DBRX), NoPE layers, logit soft-cap, attention-DP decode, context-parallel prefill (all-gather KV),
flash decoding (KV sequence-sharded inside a KV-replication group).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn as nn

from ... import ops
from ...parallel import mappings
from ...parallel.state import (Group, get_attention_dp_block_group, get_context_parallel_block_group,
                              get_context_parallel_group, get_context_parallel_tp_group,
                              get_data_parallel_attention_group, get_kv_shared_group,
                              get_tensor_model_parallel_group)
from ..gqa import GQA, GroupQueryAttention_O, GroupQueryAttention_QKV
from ..norm import L2Norm, RMSNorm


@dataclass
class AttnMeta:
    """Per-forward attention metadata shared by all layers (replaces the reference's 24-positional
    tensor ABI, model_base.py:656-718)."""
    is_prefill: bool
    position_ids: torch.Tensor                 # [B,T] rotary positions
    write_positions: torch.Tensor              # [B,T] cache slots (-1 = do not write: padding)
    seq_ids: torch.Tensor                      # [B] cache lines
    key_valid: Optional[torch.Tensor] = None   # [B,T] prefill: 1 for real tokens (padding mask)
    cos: Optional[torch.Tensor] = None
    sin: Optional[torch.Tensor] = None
    rope_cache: dict = field(default_factory=dict)   # per-rotary-module (cos, sin)
    active_mask: Optional[torch.Tensor] = None  # [B,T,M] token-tree visibility of the M slots from active_base
    active_base: Optional[torch.Tensor] = None  # [B,1] first slot covered by active_mask (default: first active slot)
    slot_mapping: Optional[torch.Tensor] = None  # paged: [B,T]
    block_table: Optional[torch.Tensor] = None   # paged: [B,max_blocks]
    prefix_len: Optional[torch.Tensor] = None    # [B] cached prefix tokens (prefix caching)
    has_prefix: bool = False
    adapter_ids: Optional[torch.Tensor] = None   # multi-LoRA
    rotary_position_ids: Optional[torch.Tensor] = None  # M-RoPE [3,B,T]
    capture: Optional[dict] = None               # tensor-capture sink
    lines: Optional[torch.Tensor] = None         # cache lines of seq_ids (computed once per forward)
    seq_hint: int = 0                            # upper bound on the live context (TKG bucket)
    extras: dict = field(default_factory=dict)    # model-specific per-forward tensors (cross-attention states, ...)


class AttentionBase(nn.Module):
    def __init__(self, config, *, hidden_size: int, num_attention_heads: int, num_key_value_heads: int,
                 head_dim: int, rotary_emb: Optional[nn.Module] = None, qkv_bias: bool = False, o_bias: bool = False,
                 sliding_window: Optional[int] = None, attention_chunk_size: Optional[int] = None,
                 qk_norm: Optional[str] = None, qk_norm_eps: float = 1e-6, clip_qkv: Optional[float] = None,
                 learned_sinks: bool = False, softmax_scale: Optional[float] = None, rope_interleaved: bool = False,
                 use_rope: bool = True, logit_softcap: Optional[float] = None, layer_idx: int = 0,
                 tensor_model_parallel_group: Optional[Group] = None, sharding_strategy: Optional[GQA] = None,
                 rms_norm_eps: float = 1e-6, device=None, qkv_input_size: Optional[int] = None):
        super().__init__()
        nc = config.neuron_config
        self.config, self.neuron_config = config, nc
        self.layer_idx = layer_idx
        self.tp_group = tensor_model_parallel_group or get_tensor_model_parallel_group()
        dtype = nc.torch_dtype
        # Odd head sizes (80, 96, 100, 112 ...): the attention kernels are specialised for 64 / 128 channels.  On the CUDA bf16 path
        # the heads are STORED zero-padded to the next kernel width (Wqkv rows / Wo columns padded at load time, KV cache that wide):
        # q.k over the padded channels adds zeros, the padded V channels stay zero, so the function is unchanged; the softmax scale and
        # the rotary tables keep the checkpoint's head size.
        self.logical_head_dim = head_dim
        rot = getattr(rotary_emb, "dim", None) if (use_rope and rotary_emb is not None) else None
        dev_is_cuda = device is not None and torch.device(device).type == "cuda"
        if (dev_is_cuda and dtype == torch.bfloat16 and head_dim not in (64, 128) and head_dim < 128 and head_dim % 2 == 0
                and qk_norm is None and attention_chunk_size is None and not logit_softcap and nc.lora_config is None
                and (rot is None or rot == head_dim) and os.environ.get("NXDI_B200_PAD_HEAD_DIM", "1") != "0"):
            head_dim = 64 if head_dim < 64 else 128
        self.hidden_size, self.head_dim = hidden_size, head_dim
        self.num_attention_heads, self.num_key_value_heads = num_attention_heads, num_key_value_heads
        sp = nc.sequence_parallel_enabled
        pad_split = not rope_interleaved
        self.qkv_proj = GroupQueryAttention_QKV(qkv_input_size or hidden_size, head_dim, num_attention_heads, num_key_value_heads,
                                                self.tp_group, dtype, qkv_bias, sharding_strategy, device,
                                                sequence_parallel_enabled=sp, src_head_dim=self.logical_head_dim, pad_split=pad_split)
        self.o_proj = GroupQueryAttention_O(hidden_size, head_dim, num_attention_heads, num_key_value_heads,
                                            self.tp_group, dtype, o_bias, sharding_strategy, device,
                                            sequence_parallel_enabled=sp, reduce_dtype=nc.rpl_reduce_dtype,
                                            src_head_dim=self.logical_head_dim, pad_split=pad_split)
        self.n_q, self.n_kv = self.qkv_proj.n_q, self.qkv_proj.n_kv
        self.rotary_emb = rotary_emb
        self.use_rope = use_rope and rotary_emb is not None
        self.rope_interleaved = rope_interleaved
        self.sliding_window = sliding_window
        self.attention_chunk_size = attention_chunk_size
        self.clip_qkv = clip_qkv
        self.softcap = logit_softcap
        self.scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(self.logical_head_dim)
        self.qk_norm = qk_norm  # None | "rms_pre_rope" | "rms_post_rope" (HunYuan; generic path) | "l2_post_rope"
        self.qk_norm_eps = qk_norm_eps
        if qk_norm in ("rms_pre_rope", "rms_post_rope"):
            self.q_layernorm = RMSNorm(head_dim, qk_norm_eps, dtype, device=device)
            self.k_layernorm = RMSNorm(head_dim, qk_norm_eps, dtype, device=device)
        elif qk_norm == "l2_post_rope":
            self.qk_l2 = L2Norm(qk_norm_eps)
        if learned_sinks:
            # one logit per local q head; sharded with the q heads
            self.sinks = nn.Parameter(torch.zeros(self.n_q, dtype=torch.float32, device=device), requires_grad=False)
            self.sinks.partition_dim = 0
            self.sinks.tp_group = self.tp_group
            plan = self.qkv_proj.plan
            self.sinks.shard_fn = lambda full, rank: torch.stack(
                [full[i] if i >= 0 else full.new_zeros(()) for i in plan.q_idx[rank]])
        else:
            self.sinks = None
        self.rms_norm_eps = rms_norm_eps
        # flash decoding: KV sequence-sharded over the ranks that replicate this rank's KV head
        self.kv_group = get_kv_shared_group() if nc.flash_decoding_enabled else None
        if self.kv_group is not None and self.kv_group.size == 1:
            self.kv_group = None
        if self.kv_group is not None and (sliding_window or attention_chunk_size or learned_sinks):
            raise NotImplementedError("flash decoding with sliding-window / chunked / sink attention")
        # attention DP (decode) and context parallelism (prefill) inside the KV-replication group: the ranks that hold the same
        # KV head split the BATCH (DP: each keeps batch/r cache lines) or the QUERY SEQUENCE (CP) instead of duplicating work
        rep = get_kv_shared_group()
        self.dp_group = rep if (nc.attention_dp_degree > 1 and rep.size == nc.attention_dp_degree) else None
        self.cp_group = rep if (nc.cp_degree > 1 and rep.size == nc.cp_degree) else None
        # Any other degree that divides tp: a block of adjacent TP ranks whose members hold DIFFERENT kv heads (or a mix).  The
        # group then also all-gathers its K/V heads (CP: per prefill; DP: the cache of a rank stores all the block's heads for ITS
        # batch rows), so that any member can serve any head of the block — the reference's `cp_degree | tp_degree` meshes
        # (attention_process_groups.py:81-110) without a second, differently sharded copy of the attention weights.
        self.cp_general = self.dp_general = False
        if nc.cp_degree > 1 and self.cp_group is None:
            self.cp_group, self.cp_general = get_context_parallel_block_group(), True
        if nc.attention_dp_degree > 1 and self.dp_group is None:
            self.dp_group, self.dp_general = get_attention_dp_block_group(), True
        if (self.cp_general or self.dp_general) and (nc.is_block_kv_layout or nc.flash_decoding_enabled):
            raise NotImplementedError("general CP / attention-DP meshes with the paged cache or flash decoding")

    # ------------------------------------------------------------------------------------
    def _rope(self, meta: AttnMeta):
        if not self.use_rope:
            return None, None
        key = id(self.rotary_emb)
        if key not in meta.rope_cache:
            pos = meta.rotary_position_ids if meta.rotary_position_ids is not None else meta.position_ids
            meta.rope_cache[key] = self.rotary_emb(pos)
        if self.head_dim != self.logical_head_dim:
            # padded heads: identity rotation (cos 1, sin 0) for the padded channel pairs
            pkey = (key, self.head_dim)
            if pkey not in meta.rope_cache:
                cos, sin = meta.rope_cache[key]
                extra = self.head_dim // 2 - cos.shape[-1]
                if self.rope_interleaved or extra <= 0:
                    padded = (torch.nn.functional.pad(cos, (0, max(extra, 0)), value=1.0), torch.nn.functional.pad(sin, (0, max(extra, 0))))
                else:
                    padded = (torch.nn.functional.pad(cos, (0, extra), value=1.0), torch.nn.functional.pad(sin, (0, extra)))
                meta.rope_cache[pkey] = padded
            return meta.rope_cache[pkey]
        return meta.rope_cache[key]

    def _simple(self) -> bool:
        """Eligible for the fused rope+norm+append kernel."""
        # (clip_qkv is applied to the projection's output before any of the fused kernels consume it)
        return (self.use_rope and not self.rope_interleaved and self.qk_norm in (None, "rms_pre_rope"))

    def forward(self, hidden: torch.Tensor, meta: AttnMeta, kv_mgr, norm_weight=None, norm_eps=None,
                norm_offset: float = 0.0, residual: Optional[torch.Tensor] = None, lora=None) -> torch.Tensor:
        """hidden [B,T,H] -> attention block output [B,T,H] (+ residual when given).
        ``norm_weight``: the layer's input RMSNorm, fused into the QKV projection."""
        B, T, _ = hidden.shape
        if self.qkv_proj.sequence_parallel_enabled:
            T = T * self.tp_group.size            # hidden is the local sequence shard; qkv_proj gathers it
        D, nq, nkv = self.head_dim, self.n_q, self.n_kv
        qkv = self.qkv_proj(hidden, norm_weight, norm_eps if norm_eps is not None else self.rms_norm_eps, norm_offset)
        if lora is not None and lora.has("qkv_proj"):
            xn = ops.rmsnorm(hidden, norm_weight, norm_eps if norm_eps is not None else self.rms_norm_eps, norm_offset) \
                if norm_weight is not None else hidden
            qkv = qkv + lora("qkv_proj", xn, meta.adapter_ids)
        if self.clip_qkv is not None:
            qkv = qkv.clamp(-self.clip_qkv, self.clip_qkv)
        cos, sin = self._rope(meta)
        paged = meta.slot_mapping is not None
        if self.kv_group is not None:
            return self._forward_flash_decoding(qkv, meta, kv_mgr, cos, sin, residual, B, T)
        if (self.dp_group is not None and (not meta.is_prefill or self.dp_general)) or \
                (self.cp_group is not None and meta.is_prefill and not meta.has_prefix and T % self.cp_group.size == 0):
            return self._forward_group_parallel(qkv, meta, kv_mgr, cos, sin, residual, B, T)
        k_cache, v_cache = kv_mgr.get_kv_by_layer_id(self.layer_idx)
        if meta.lines is None:
            meta.lines = meta.seq_ids if paged else kv_mgr.lines_for(meta.seq_ids)
        lines = meta.lines
        # rolling (window-sized) caches: slots are positions modulo the window, the horizon saturates at W-1 and every
        # resident slot is inside the window (modules/kvcache/gpt_oss_kv_cache_manager.py)
        wpos, hpos, window, hint = meta.write_positions, meta.position_ids, self.sliding_window, meta.seq_hint
        roll = kv_mgr.rolling_window(self.layer_idx) if hasattr(kv_mgr, "rolling_window") else None
        if roll:
            if paged or meta.has_prefix or meta.active_mask is not None:
                raise NotImplementedError("rolling sliding-window cache with paged KV / cached prefixes / token trees")
            wpos, hpos = kv_mgr.rolling_positions(roll, meta.write_positions, meta.position_ids, meta.is_prefill)
            if not meta.is_prefill:
                window, hint = None, min(hint, roll) if hint else hint
        fused = (self._simple() and not paged and qkv.is_cuda and k_cache.dtype == qkv.dtype
                 and qkv.dtype in (torch.bfloat16, torch.float16)
                 and (not meta.is_prefill or not meta.has_prefix))
        if fused:
            qn = self.q_layernorm.weight if self.qk_norm == "rms_pre_rope" else None
            kn = self.k_layernorm.weight if self.qk_norm == "rms_pre_rope" else None
            k = v = None
            if meta.is_prefill:
                # prefill attention consumes the fresh k/v directly (no cache read): split here.  Plain bf16 cache: split + q/k norm +
                # RoPE + cache write in one kernel; managers with their own update logic (quantised, rolling, hybrid) keep theirs
                qkvs = None
                if type(kv_mgr).__name__ == "KVCacheManager" and not roll and getattr(kv_mgr, "k_scale", None) is None \
                        and meta.capture is None:
                    qkvs = ops.rope_kv_split_append(qkv, cos, sin, k_cache, v_cache, lines, wpos, nq, nkv, D, qn, kn, self.qk_norm_eps)
                if qkvs is not None:
                    q, k, v = qkvs
                else:
                    q, k, v = self._split_norm_rope(qkv, B, T, cos, sin, meta)
                    kv_mgr.update(self.layer_idx, k, v, meta.seq_ids, wpos, lines)
            elif (meta.active_mask is None and self.attention_chunk_size is None and not self.softcap
                  and getattr(kv_mgr, "k_scale", None) is None and meta.capture is None):
                o = ops.rope_attention_decode(qkv, cos, sin, k_cache, v_cache, lines, wpos, hpos,
                                              nq, nkv, D, self.scale, window, self.sinks, qn, kn,
                                              self.qk_norm_eps, seq_hint=hint)
                return self._finish(o.reshape(B, T, nq * D), residual, lora, meta)
            else:
                q = ops.rope_kv_append(qkv, cos, sin, k_cache, v_cache, lines, wpos, nq, nkv, D,
                                       False, qn, kn, self.qk_norm_eps)
        else:
            q, k, v = self._split_norm_rope(qkv, B, T, cos, sin, meta)
            if paged:
                kv_mgr.update(self.layer_idx, k, v, meta.slot_mapping)
            else:
                kv_mgr.update(self.layer_idx, k, v, meta.seq_ids, wpos, lines)
        if meta.capture is not None:
            meta.capture[f"layers.{self.layer_idx}.self_attn.q"] = q

        if meta.is_prefill and not meta.has_prefix and meta.extras.get("bidir_group_ids") is not None:
            o = self._prefill_bidirectional_groups(q, k, v, meta)
        elif meta.is_prefill and not meta.has_prefix:
            right = self._arange_pos(meta)
            # right padding: pad keys sit after every real token, causality alone hides them
            o = ops.attention_prefill(q, k, v, self.scale, True, self.sliding_window, self.attention_chunk_size,
                                      None if right else meta.key_valid, None if right else meta.position_ids,
                                      self.sinks, self.softcap)
        elif paged:
            if self.attention_chunk_size is not None or meta.active_mask is not None:
                raise NotImplementedError("chunked attention / token trees with the paged cache")
            o = ops.paged_attention_decode(q, k_cache, v_cache, meta.block_table, meta.position_ids,
                                           self.scale, self.sliding_window, self.sinks)
        else:
            ks = getattr(kv_mgr, "k_scale", None)
            vs = getattr(kv_mgr, "v_scale", None)
            # token trees: visibility is defined on cache slots (node index), rotary positions are depths
            vis_pos = meta.write_positions if meta.active_mask is not None else hpos
            o = ops.attention_decode(q, k_cache, v_cache, lines, vis_pos, self.scale, window,
                                     self.attention_chunk_size, self.sinks, meta.active_mask, self.softcap, ks, vs,
                                     seq_hint=hint, active_base=meta.active_base)
        return self._finish(o.reshape(B, T, nq * D), residual, lora, meta)

    def _prefill_bidirectional_groups(self, q, k, v, meta):
        """Causal (+ window / chunk) prefill where tokens sharing a non-negative group id (the soft tokens of one image) also see
        each other in BOTH directions — Gemma-3's image attention.  Masked fp32 path; image prompts are prefill-only work."""
        grp = meta.extras["bidir_group_ids"].to(q.device)
        B, T = q.shape[:2]
        if grp.shape[1] < T:                                  # the runner padded the prompt to its bucket
            grp = torch.nn.functional.pad(grp, (0, T - grp.shape[1]), value=-1)
        right = self._arange_pos(meta)
        pos = torch.arange(T, device=q.device).unsqueeze(0).expand(B, T) if right else meta.position_ids
        mask = ops.ref.build_mask(pos, T, self.sliding_window, self.attention_chunk_size, None if right else meta.key_valid)
        same = (grp.unsqueeze(2) == grp.unsqueeze(1)) & (grp.unsqueeze(2) >= 0)
        o = ops.ref.attention_with_mask(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), mask | same.unsqueeze(1), self.scale,
                                        self.sinks, self.softcap)
        return o.transpose(1, 2)

    def _finish(self, o, residual, lora, meta):
        out = self.o_proj(o, residual)
        if lora is not None and lora.has("o_proj"):
            from ...parallel import mappings as _m
            out = out + _m.all_reduce(lora("o_proj", o, meta.adapter_ids), self.tp_group)
        return out

    def chain_eligible(self, kv_mgr, dtype) -> bool:
        """This layer's decode step can run as [rope+append kernel, attention kernel] around the persistent GEMV chain."""
        k_cache, _ = kv_mgr.get_kv_by_layer_id(self.layer_idx)
        return (self._simple() and self.clip_qkv is None and self.kv_group is None and k_cache.dtype == dtype and self.head_dim in (64, 128)
                and getattr(self.qkv_proj, "scale", None) is None and getattr(self.o_proj, "scale", None) is None
                and not self.qkv_proj.sequence_parallel_enabled and self.attention_chunk_size is None and not self.softcap)

    def decode_core(self, qkv, meta: AttnMeta, kv_mgr, B: int, T: int) -> torch.Tensor:
        """qkv [B,T,(nq+2nkv)D] (already projected) -> attention output [B,T,nq*D]: fused q/k norm + RoPE + cache append,
        then split-KV flash decode."""
        D, nq, nkv = self.head_dim, self.n_q, self.n_kv
        cos, sin = self._rope(meta)
        k_cache, v_cache = kv_mgr.get_kv_by_layer_id(self.layer_idx)
        if meta.lines is None:
            meta.lines = kv_mgr.lines_for(meta.seq_ids)
        qn = self.q_layernorm.weight if self.qk_norm == "rms_pre_rope" else None
        kn = self.k_layernorm.weight if self.qk_norm == "rms_pre_rope" else None
        o = ops.rope_attention_decode(qkv, cos, sin, k_cache, v_cache, meta.lines, meta.write_positions, meta.position_ids,
                                      nq, nkv, D, self.scale, self.sliding_window, self.sinks, qn, kn, self.qk_norm_eps,
                                      seq_hint=meta.seq_hint)
        return o.reshape(B, T, nq * D)

    def _forward_group_parallel(self, qkv, meta, kv_mgr, cos, sin, residual, B, T):
        """Attention DP (decode) / CP (prefill) among the r ranks that replicate this rank's KV head.
        Both start by all-gathering the q heads of the group, so any member can serve any of the group's heads:
          * DP decode: rank j owns the cache lines of batch rows ``seq_id in [j*n, (j+1)*n)`` (DataParallelKVCacheManager);
            it attends ALL group heads for ITS rows, zeros elsewhere, and one all-reduce inside the group assembles the rows;
          * CP prefill: rank j attends all group heads for the j-th slice of the query sequence against the full fresh K/V
            (replicated anyway), slices are all-gathered along the sequence.
        reference: attention_base.py:1709-1727 (DP) / :603-630 (CP, all-gather-KV); there both need a second, tp/dp- or
        tp/cp-sharded copy of the attention weights — here the existing head sharding is reused as is."""
        D, nq = self.head_dim, self.n_q
        q, k, v = self._split_norm_rope(qkv, B, T, cos, sin, meta)
        lines = kv_mgr.lines_for(meta.seq_ids)
        ka = va = None
        if self.dp_general:
            # general attention DP: this rank's cache keeps ALL kv heads of its block for the batch rows it owns
            ka = mappings.all_gather(k.contiguous(), 2, self.dp_group)
            va = mappings.all_gather(v.contiguous(), 2, self.dp_group)
            kv_mgr.update(self.layer_idx, ka, va, meta.seq_ids, meta.write_positions, lines)
        else:
            kv_mgr.update(self.layer_idx, k, v, meta.seq_ids, meta.write_positions, lines)     # non-owned rows -> garbage line

        def gathered(t, g):
            return None if t is None else mappings.all_gather(t.contiguous(), 0, g)
        if meta.is_prefill:
            g = self.cp_group
            if g is None or meta.has_prefix or T % g.size != 0:
                # (general attention DP without CP: the prefill itself is ordinary attention over the local heads)
                right = self._arange_pos(meta)
                o = ops.attention_prefill(q, k, v, self.scale, True, self.sliding_window, self.attention_chunk_size,
                                          None if right else meta.key_valid, None if right else meta.position_ids, self.sinks, self.softcap)
                return self.o_proj(o.reshape(B, T, nq * D), residual)
            n = T // g.size
            # contiguous slices by default; strided (rank j takes positions j, j + r, ...) balances the causal work between the
            # ranks — the reference's strided_context_parallel_kernel_enabled (attention_base.py:546-561)
            strided = bool(getattr(self.neuron_config, "strided_context_parallel_kernel_enabled", False))
            sl = slice(g.rank, None, g.size) if strided else slice(g.rank * n, (g.rank + 1) * n)
            qa = mappings.all_gather(q.contiguous(), 2, g)[:, sl]                                 # [B, T/r, r*nq, D]
            qpos = meta.position_ids[:, sl]
            kk, vv = k, v
            if self.cp_general:      # members hold different kv heads: gather them too (rank-major, like the q heads)
                same = self.dp_general and self.dp_group is g
                kk = ka if same else mappings.all_gather(k.contiguous(), 2, g)
                vv = va if same else mappings.all_gather(v.contiguous(), 2, g)
            o = ops.attention_prefill(qa.contiguous(), kk, vv, self.scale, True, self.sliding_window, self.attention_chunk_size,
                                      meta.key_valid, qpos, gathered(self.sinks, g), self.softcap)
            o = mappings.all_gather(o.contiguous(), 1, g)                                          # [B, T, r*nq, D]
            if strided:      # rank-major [r][T/r] -> interleaved sequence order
                o = o.view(B, g.size, n, o.shape[2], D).transpose(1, 2).reshape(B, T, o.shape[2], D)
        else:
            g = self.dp_group
            k_cache, v_cache = kv_mgr.get_kv_by_layer_id(self.layer_idx)
            qa = mappings.all_gather(q.contiguous(), 2, g)
            o = ops.attention_decode(qa, k_cache, v_cache, lines, meta.position_ids, self.scale, self.sliding_window,
                                     self.attention_chunk_size, gathered(self.sinks, g), None, self.softcap, seq_hint=meta.seq_hint)
            owned = (lines < kv_mgr.num_lines).view(B, 1, 1, 1)
            o = torch.where(owned, o, torch.zeros_like(o))
            o = mappings.all_reduce(o.float(), g).to(q.dtype)
        o = o[:, :, g.rank * nq:(g.rank + 1) * nq]
        return self.o_proj(o.reshape(B, T, nq * D), residual)

    def _forward_flash_decoding(self, qkv, meta, kv_mgr, cos, sin, residual, B, T):
        from .. import flashdecode as fd
        g = self.kv_group
        D, nq = self.head_dim, self.n_q
        q, k, v = self._split_norm_rope(qkv, B, T, cos, sin, meta)
        lines = kv_mgr.lines_for(meta.seq_ids)
        kv_mgr.update(self.layer_idx, k, v, meta.seq_ids, fd.local_slots(meta.write_positions, g.rank, g.size), lines)
        if meta.is_prefill and not meta.has_prefix:
            right = self._arange_pos(meta)
            o = ops.attention_prefill(q, k, v, self.scale, True, None, None, None if right else meta.key_valid,
                                      None if right else meta.position_ids, None, self.softcap)
        else:
            k_cache, v_cache = kv_mgr.get_kv_by_layer_id(self.layer_idx)
            qa = mappings.all_gather(q.contiguous(), 2, g)                       # [B,T,r*nq,D]: all q heads of the KV group
            li = lines.long().clamp(0, k_cache.shape[0] - 1)
            po, pm, pl = fd.partial_attention(qa, k_cache[li], v_cache[li], fd.local_horizon(meta.position_ids, g.rank, g.size),
                                              self.scale)
            o = fd.combine(po, pm, pl, g)[:, :, g.rank * nq:(g.rank + 1) * nq].to(q.dtype)
        return self.o_proj(o.reshape(B, T, nq * D), residual)

    def _arange_pos(self, meta: AttnMeta) -> bool:
        """Prefill kernels assume q position == token index (right padding).  Left padding / offset
        starts take the masked path."""
        return self.neuron_config.padding_side == "right" and not getattr(meta, "offset_positions", False)

    def _split_norm_rope(self, qkv, B, T, cos, sin, meta=None):
        D, nq, nkv = self.head_dim, self.n_q, self.n_kv
        q, k, v = qkv.reshape(B, T, nq + 2 * nkv, D).split([nq, nkv, nkv], dim=2)
        if self.qk_norm == "rms_pre_rope":
            q = self.q_layernorm(q)
            k = self.k_layernorm(k)
        if cos is not None:
            q = ops.apply_rope(q, cos, sin, self.rope_interleaved)
            k = ops.apply_rope(k, cos, sin, self.rope_interleaved)
        if self.qk_norm == "l2_post_rope":
            q = self.qk_l2(q)
            k = self.qk_l2(k)
        elif self.qk_norm == "rms_post_rope":
            q = self.q_layernorm(q)
            k = self.k_layernorm(k)
        return q, k, v
