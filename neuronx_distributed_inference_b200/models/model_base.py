"""Device-side decoder model: embed -> layer loop (KV cache in place) -> last-token gather ->
lm_head -> on-device sampling.

Role of reference ``NeuronBaseModel`` (models/model_base.py:70-1596).  Differences by design:
* no tracing: this is an eager ``nn.Module`` replayed through per-bucket CUDA graphs
  (runtime/runner.py); masks are never materialised — kernels take positions / lengths;
* one typed :class:`AttnMeta` per forward instead of 24 positional tensors;
* KV cache updates are in place inside the attention block (no aliasing, no deferred write).
Subclasses implement ``setup_attr_for_model`` + ``init_model`` exactly like the reference hooks
(model_base.py:120-141).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import ops
from ..modules.attention import AttnMeta
from ..modules.kvcache import BlockKVCacheManager, DataParallelKVCacheManager, KVCacheManager
from ..modules.sampling import Sampler, mask_padded_logits
from ..parallel import mappings
from ..parallel.state import get_tensor_model_parallel_group, get_tp_group


@dataclass
class ModelOutput:
    """What one forward returns (device tensors)."""
    tokens: Optional[torch.Tensor] = None          # [B] or [B,T] sampled ids (on-device sampling)
    logits: Optional[torch.Tensor] = None          # [B,T_out,V] (full vocab) when requested
    hidden_states: Optional[torch.Tensor] = None   # [B,T_out,H] last hidden (EAGLE / capture)
    captured: Optional[Dict[str, torch.Tensor]] = None
    extras: dict = field(default_factory=dict)


class DecoderLayer(nn.Module):
    """Pre-norm transformer block with both norms fused into the following projection and both
    residual adds fused into the preceding row-parallel projection."""

    def __init__(self, attn: nn.Module, mlp: nn.Module, input_layernorm: nn.Module,
                 post_attention_layernorm: nn.Module, layer_idx: int = 0, mlp_is_moe: bool = False):
        super().__init__()
        self.self_attn = attn
        self.mlp = mlp
        self.input_layernorm = input_layernorm
        self.post_attention_layernorm = post_attention_layernorm
        self.layer_idx = layer_idx
        self.mlp_is_moe = mlp_is_moe

    def forward(self, h: torch.Tensor, meta: AttnMeta, kv_mgr, lora=None) -> torch.Tensor:
        n1, n2 = self.input_layernorm, self.post_attention_layernorm
        h = self.self_attn(h, meta, kv_mgr, norm_weight=n1.weight, norm_eps=n1.variance_epsilon,
                           norm_offset=n1.offset, residual=h, lora=lora)
        if meta.capture is not None:
            meta.capture[f"layers.{self.layer_idx}.attn_out"] = h
        if lora is not None and not self.mlp_is_moe:
            h = self.mlp(h, norm_weight=n2.weight, norm_eps=n2.variance_epsilon, norm_offset=n2.offset, residual=h, lora=lora,
                         adapter_ids=meta.adapter_ids)
        else:
            h = self.mlp(h, norm_weight=n2.weight, norm_eps=n2.variance_epsilon, norm_offset=n2.offset, residual=h)
        if meta.capture is not None:
            meta.capture[f"layers.{self.layer_idx}.out"] = h
        return h


class NeuronBaseModel(nn.Module):
    """Generic causal decoder.  Subclass hooks: ``setup_attr_for_model(config)`` (set
    ``tp_degree, hidden_size, num_attention_heads, num_key_value_heads, max_batch_size,
    buckets``...) and ``init_model(config)`` (create ``embed_tokens, layers, norm, lm_head``)."""

    def __init__(self, config, device=None):
        super().__init__()
        self.config = config
        self.neuron_config = nc = config.neuron_config
        self.device_ = device
        self.tp_group = get_tp_group(config)
        self.vocab_size = getattr(config, "vocab_size", None)
        self.padding_side = nc.padding_side
        self.on_device_sampling = nc.on_device_sampling_config is not None
        self.setup_attr_for_model(config)
        self.init_model(config)
        self.init_inference_optimization(config)
        # CUDA-graph capture needs every op of the decode step to be sync-free: true for the hand-written kernel path
        # (bf16, head_dim 64/128); the PyTorch composite fallbacks (odd head dims, fp8 KV, expert dispatch) are not.
        self.graph_safe = bool(getattr(type(self), "graph_safe", True)) and (self._kernels_cover_decode() or self._fallback_is_sync_free())
        if not self.graph_safe and getattr(type(self), "moe_decode_graph_safe", False):
            # MoE families whose routed experts run through the moe_decode kernels (no host-side dispatch)
            self.graph_safe = self._kernels_cover_decode() and self._moe_kernels_cover_decode()

    def _fallback_is_sync_free(self) -> bool:
        """Dense decoder on the contiguous cache with a garbage line: the PyTorch composite path has no host sync either."""
        nc = self.neuron_config
        return (not nc.is_block_kv_layout and not nc.kv_cache_quant and getattr(self.kv_mgr, "garbage", 0) == 1
                and not any(getattr(l, "mlp_is_moe", False) for l in self.layers)
                # head rows that are not 16-byte multiples take the PyTorch cache append (boolean indexing: a host sync)
                and all((getattr(getattr(l, "self_attn", None), "head_dim", 8) * 2) % 16 == 0 for l in self.layers))

    def _moe_kernels_cover_decode(self) -> bool:
        if self.device_ is None or torch.device(self.device_).type != "cuda":
            return False
        for layer in self.layers:
            if not getattr(layer, "mlp_is_moe", False):
                continue
            moe = layer.mlp
            ex = getattr(moe, "expert_mlps", None)
            if ex is None or ex.gate_up_proj.dtype != torch.bfloat16:
                return False
            plain = (not getattr(moe, "early_affinity_modulation", False) and ex.act == "silu_mul" and ex.act_fn is None
                     and ex.gate_up_bias is None and ex.down_bias is None)        # moe_decode kernels (T <= 8)
            kact = ex.act if ex.act_fn is None else getattr(ex.act_fn, "kernel_act", None)
            grouped = (kact in ops._MOE_ACTS and ex.gate_up_proj.shape[2] % 64 == 0 and ex.down_proj.shape[2] % 64 == 0
                       and ex.gate_up_proj.shape[0] <= 512 and getattr(ex, "gated", True))   # grouped tcgen05 GEMMs (any T)
            if not (plain or grouped):
                return False
        return True

    def _moe_kernels_cover_prefill(self) -> bool:
        """Routed experts of a prefill-sized batch run through the device-side permutation + grouped tcgen05 GEMMs
        (ops.moe_experts -> csrc/moe_grouped.cu): static shapes, no host sync -> context encoding may be graph-captured."""
        if not self._moe_kernels_cover_decode():
            return False
        for layer in self.layers:
            ex = getattr(getattr(layer, "mlp", None), "expert_mlps", None) if getattr(layer, "mlp_is_moe", False) else None
            if ex is not None and (ex.gate_up_proj.shape[2] % 64 != 0 or ex.down_proj.shape[2] % 64 != 0 or ex.gate_up_proj.shape[0] > 512):
                return False
        return True

    def _kernels_cover_decode(self) -> bool:
        nc = self.neuron_config
        if nc.torch_dtype != torch.bfloat16 or nc.kv_cache_quant:
            return False
        for layer in self.layers:
            attn = getattr(layer, "self_attn", None)
            if attn is None:
                continue
            if getattr(attn, "head_dim", 128) not in (64, 128):
                return False
            if getattr(attn, "attention_chunk_size", None) is not None or getattr(attn, "softcap", None):
                return False
            if hasattr(attn, "_simple") and not attn._simple():
                return False
        return True

    # hooks ---------------------------------------------------------------------------------
    def setup_attr_for_model(self, config):
        raise NotImplementedError

    def init_model(self, config):
        raise NotImplementedError

    # ------------------------------------------------------------------------------------
    def kv_heads_per_rank(self) -> int:
        return self.layers[0].self_attn.n_kv

    def kv_head_dim(self) -> int:
        return self.layers[0].self_attn.head_dim

    def init_inference_optimization(self, config):
        nc = self.neuron_config
        if self.on_device_sampling:
            self.sampler = Sampler(nc, self.tp_group, vocab_shard=self.lm_head_is_sharded())
        n_layers = len(self.layers)
        dtype = nc.attention_dtype or nc.torch_dtype
        if nc.is_block_kv_layout:
            self.kv_mgr = BlockKVCacheManager(n_layers, self.kv_heads_per_rank(), self.kv_head_dim(),
                                              nc.pa_num_blocks, nc.pa_block_size, dtype, self.device_)
        else:
            lines = nc.kv_cache_batch_size + nc.kv_cache_padding_size
            kw = dict(num_layers=n_layers, num_kv_heads=self.kv_heads_per_rank(), head_dim=self.kv_head_dim(),
                      max_len=self._kv_len(config),
                      num_lines=lines, dtype=dtype, device=self.device_,
                      quant_config=nc.kv_quant_config if nc.kv_cache_quant else None)
            if nc.attention_dp_degree > 1:
                # attention DP lives inside the KV-replication group (ranks that would otherwise hold identical KV); any other
                # degree: a block of adjacent TP ranks, each keeping ALL the block's kv heads for its share of the batch rows
                attn0 = self.layers[0].self_attn
                g = attn0.dp_group
                if getattr(attn0, "dp_general", False):
                    kw["num_kv_heads"] = self.kv_heads_per_rank() * g.size
                self.kv_mgr = DataParallelKVCacheManager(dp_rank=g.rank, dp_size=g.size, **kw)
            elif nc.rolling_sliding_window_cache and any(self._layer_windows()):
                from ..modules.kvcache.gpt_oss_kv_cache_manager import HybridKVCacheManager
                kw.pop("num_layers")
                self.kv_mgr = HybridKVCacheManager(self._layer_windows(), **kw)
            else:
                self.kv_mgr = KVCacheManager(**kw)

    def _layer_windows(self):
        """Per-layer sliding window (None for full-attention layers)."""
        return [getattr(getattr(layer, "self_attn", None), "sliding_window", None) for layer in self.layers]

    def _kv_len(self, config) -> int:
        nc = self.neuron_config
        n = nc.max_length + self._speculation_slack()
        r = getattr(config, "num_cores_per_group", 1) if nc.flash_decoding_enabled else 1
        return -(-n // r)          # flash decoding: each rank of a KV group keeps 1/r of the positions

    def _speculation_slack(self) -> int:
        """Extra cache slots past ``max_length``: a verify step writes every candidate (chain of k, or all tree nodes)
        before acceptance."""
        nc = self.neuron_config
        extra = nc.speculation_length or 0
        if nc.token_tree_config is not None or nc.is_medusa:
            from ..modules.eagle.token_tree import TokenTree
            if nc.token_tree_config is not None:
                extra = max(extra, TokenTree(nc.token_tree_config).num_nodes)
            if nc.is_medusa:
                from ..generation.medusa import DEFAULT_MEDUSA_TREE
                extra = max(extra, TokenTree(nc.medusa_tree or DEFAULT_MEDUSA_TREE).num_nodes)
        return extra

    def lm_head_is_sharded(self) -> bool:
        return self.tp_group.size > 1 and not getattr(self.lm_head, "gather_output", True)

    # ------------------------------------------------------------------------------------
    def embed(self, input_ids, inputs_embeds=None, vision_embeddings=None, vision_mask=None):
        h = inputs_embeds if inputs_embeds is not None else self.embed_tokens(input_ids)
        if vision_embeddings is not None and vision_mask is not None:
            h = self.encode_vision_to_input(h, vision_embeddings, vision_mask)
        scale = getattr(self, "embed_scale", None)
        if scale is not None:
            h = h * torch.tensor(scale, dtype=h.dtype, device=h.device)
        return h

    def encode_vision_to_input(self, h, vision_embeddings, vision_mask):
        """Scatter image embeddings into the token embeddings at masked positions
        (reference model_base.py:1193-1200)."""
        m = vision_mask.bool()
        if m.dim() == 3:
            m = m.squeeze(-1)
        h = h.clone()
        h[m] = vision_embeddings.reshape(-1, h.shape[-1])[: int(m.sum())].to(h.dtype)
        return h

    def build_meta(self, input_ids, attention_mask, position_ids, seq_ids, is_prefill, **kw) -> AttnMeta:
        B, T = input_ids.shape[:2]
        dev = input_ids.device
        if seq_ids is None:
            seq_ids = torch.arange(B, device=dev, dtype=torch.int32)
        if position_ids is None:
            position_ids = torch.arange(T, device=dev).unsqueeze(0).expand(B, T)
        key_valid = None
        has_prefix = bool(kw.get("has_prefix", False))
        if is_prefill:
            if attention_mask is not None and attention_mask.shape[-1] == T:
                key_valid = attention_mask.to(torch.bool)
            if key_valid is not None:
                write = torch.where(key_valid, position_ids, torch.full_like(position_ids, -1))
            else:
                write = position_ids
        else:
            write = position_ids
        if kw.get("write_positions") is not None:     # token trees: rotary position = depth, cache slot = node index
            write = kw["write_positions"]
        meta = AttnMeta(is_prefill=is_prefill, position_ids=position_ids.to(torch.int32),
                        write_positions=write.to(torch.int32), seq_ids=seq_ids.to(torch.int32), key_valid=key_valid,
                        active_mask=kw.get("active_mask"), active_base=kw.get("active_base"), slot_mapping=kw.get("slot_mapping"),
                        block_table=kw.get("block_table"), has_prefix=has_prefix,
                        adapter_ids=kw.get("adapter_ids"), rotary_position_ids=kw.get("rotary_position_ids"),
                        capture={} if kw.get("capture") else None, seq_hint=int(kw.get("seq_hint") or 0))
        for k in getattr(self, "meta_extra_keys", ()):
            if kw.get(k) is not None:
                meta.extras[k] = kw[k]
        if self.padding_side != "right" or kw.get("offset_positions"):
            meta.offset_positions = True
        return meta

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, seq_ids: Optional[torch.Tensor] = None,
                sampling_params: Optional[torch.Tensor] = None, *, is_prefill: Optional[bool] = None,
                prev_hidden: Optional[torch.Tensor] = None, inputs_embeds=None, vision_embeddings=None,
                vision_mask=None, rand=None, output_logits: Optional[bool] = None, output_hidden: bool = False,
                all_positions: bool = False, all_hidden: bool = False, **kw) -> ModelOutput:
        nc = self.neuron_config
        B, T = input_ids.shape[:2]
        if is_prefill is None:
            is_prefill = T > 1 and T != nc.speculation_length and T != nc.medusa_speculation_length
        meta = self.build_meta(input_ids, attention_mask, position_ids, seq_ids, is_prefill, **kw)
        if self.tp_group.symm is not None:
            self.tp_group.symm.begin_step()     # fresh tags for this forward's in-kernel collectives (parallel/symm.py)
        h = self.embed(input_ids, inputs_embeds, vision_embeddings, vision_mask)
        sp = self._set_sequence_parallel(is_prefill and T % self.tp_group.size == 0 and T >= self.tp_group.size)
        if sp:
            # sequence parallel prefill (reference model_base.py:1471-1583): the residual stream is sharded along the
            # sequence between blocks; column-parallel layers all-gather it, row-parallel layers reduce-scatter back
            n = T // self.tp_group.size
            h = h[:, self.tp_group.rank * n:(self.tp_group.rank + 1) * n].contiguous()
        if prev_hidden is not None and hasattr(self, "fuse_prev_hidden"):
            h = self.fuse_prev_hidden(h, prev_hidden)
        lora = getattr(self, "lora", None) if meta.adapter_ids is not None else None
        aux_layers = getattr(self, "aux_hidden_layers", None) if output_hidden else None
        aux = []
        mega = False
        if not is_prefill and lora is None and aux_layers is None and prev_hidden is None:
            from ..runtime import decode_step
            if decode_step.eligible(self, h, meta, kw):
                # every layer of the decode step in ONE persistent kernel (csrc/decode_step.cu)
                h = decode_step.run_layers(self, h, meta)
                mega = True
        for i, layer in enumerate(() if mega else self.layers):
            h = layer(h, meta, self.kv_mgr, lora=lora.for_layer(i)) if lora is not None else layer(h, meta, self.kv_mgr)
            if aux_layers is not None and i in aux_layers:
                aux.append(h)
            ds = kw.get("deepstack_embeds")
            if ds is not None and is_prefill and i < len(ds) and vision_mask is not None:
                # Qwen3-VL deepstack: intermediate vision features are added to the residual stream of the first layers
                vm = vision_mask.bool().squeeze(-1) if vision_mask.dim() == 3 else vision_mask.bool()
                h = h.clone()
                h[vm] = h[vm] + ds[i].to(h.dtype)[: int(vm.sum())]
        if sp:
            h = mappings.all_gather(h.contiguous(), 1, self.tp_group)
        # ---- last-token gather (prefill) ------------------------------------------------------
        if is_prefill and not all_positions:
            if meta.key_valid is not None and self.padding_side == "right":
                last = meta.key_valid.sum(-1).clamp_min(1) - 1
            elif self.padding_side == "right" and attention_mask is None:
                last = torch.full((B,), T - 1, device=h.device, dtype=torch.long)
            else:
                last = meta.position_ids.long().argmax(-1)  # reference model_base.py:974-976
            h_out = h[torch.arange(B, device=h.device), last.long()].unsqueeze(1)
        else:
            h_out = h
        out = ModelOutput()
        if output_hidden:
            # EAGLE-3 targets export the concatenation of a low / middle / high layer (raw residual stream), EAGLE-1/2 the
            # final normalised state; ``all_hidden`` keeps every prompt position (draft prefill) while logits stay last-only
            src = h if all_hidden else h_out
            if aux:
                cat = torch.cat(aux, -1)
                out.hidden_states = cat if (all_hidden or not (is_prefill and not all_positions)) else \
                    cat[torch.arange(B, device=h.device), last.long()].unsqueeze(1)
            else:
                out.hidden_states = self.final_hidden(src)
        logits = self.compute_logits(h_out)
        want_logits = nc.output_logits if output_logits is None else output_logits
        if self.on_device_sampling:
            Bo, To, V = logits.shape
            toks = self.sampler(logits.reshape(Bo * To, V),
                                None if sampling_params is None else sampling_params.repeat_interleave(To, 0),
                                rand if (rand is None or To == 1 or rand.shape[0] == Bo * To) else rand.repeat_interleave(To, 0))
            out.tokens = toks.view(Bo, To) if To > 1 else toks.view(Bo)
        if want_logits or not self.on_device_sampling:
            out.logits = self.gather_logits(logits)
        out.captured = meta.capture
        if getattr(self, "medusa_heads", None) is not None:
            out.extras["medusa_logits"] = self.medusa_logits(h_out)
        return out

    def medusa_logits(self, h_out):
        """[num_heads, B, T, V]: every Medusa head = ResBlock(s) + its own vocab projection on the final hidden state
        (reference model_base.py:478-508, modeling_llama.py:1172-1187)."""
        hn = self.final_hidden(h_out)
        return torch.stack([self.gather_logits(head(hn)) for head in self.medusa_heads], 0)

    def _set_sequence_parallel(self, on: bool) -> bool:
        """SP is a prefill-only layout; the parallel layers carry a static flag in the reference, here it is switched per
        forward so that the same modules serve decode (T=1, nothing to shard)."""
        if not self.neuron_config.sequence_parallel_enabled or self.tp_group.size == 1:
            return False
        if not hasattr(self, "_sp_modules"):
            self._sp_modules = [m for m in self.layers.modules() if hasattr(m, "sequence_parallel_enabled")]
        for m in self._sp_modules:
            m.sequence_parallel_enabled = on
            if getattr(m, "sequence_dimension", None) is None:
                m.sequence_dimension = 1
        return on

    def final_hidden(self, h):
        return self.norm(h)

    def compute_logits(self, h) -> torch.Tensor:
        """Final norm fused into the lm_head GEMM; returns the *local* vocab shard [B,T,V/tp]."""
        n = self.norm
        logits = self.lm_head(h, norm_weight=n.weight, norm_eps=n.variance_epsilon, norm_offset=n.offset)
        soft = getattr(self, "final_logit_softcap", None)
        if soft:
            logits = torch.tanh(logits.float() / soft) * soft
        pad = getattr(self.lm_head, "pad_size", 0)
        if pad and self.lm_head_is_sharded():
            logits = mask_padded_logits(logits, self.tp_group.rank, self.tp_group.size, pad)
        return logits

    def gather_logits(self, logits):
        if self.lm_head_is_sharded():
            logits = mappings.all_gather(logits, -1, self.tp_group)
            pad = getattr(self.lm_head, "pad_size", 0)
            if pad:
                logits = logits[..., : logits.shape[-1] - pad]
        return logits.float()

    def reset(self):
        self.kv_mgr.reset()
