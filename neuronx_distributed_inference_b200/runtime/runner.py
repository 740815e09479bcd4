"""Per-sub-model runner: bucket selection, padding, batch splitting, pinned H2D staging and
CUDA-graph replay.

Role of reference ``ModelWrapper`` (models/model_wrapper.py:50-1440) + ``ModelBuilder``/``NxDModel``
routing.  What remains of it on B200:
* a *bucket* is a CUDA-graph key ``(batch bucket, seq bucket, n_active)``; kernels read real
  lengths from device memory so no attention masks are padded or shipped;
* batch rows are never sorted by ``seq_ids`` (kernels index cache lines directly); short batches
  are padded with ``seq_id=-1`` rows that write to the garbage line (reference pads by repeating row
  0 and sorting, model_wrapper.py:520-703);
* inputs go through pinned staging buffers with async H2D copies; outputs come back through a pinned
  D2H buffer so a step costs one stream sync;
* async mode (reference modules/async_execution.py): the graph itself feeds the sampled token and
  ``position+1`` into the next step's static inputs, so steps can be enqueued back-to-back and the
  host reads tokens one step late.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from ..modules import autobucketing
from ..models.model_base import ModelOutput

logger = logging.getLogger("b200infer")


def _to_dev(t, device, dtype=None, non_blocking=True):
    if t is None:
        return None
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device != device:
        if t.device.type == "cpu" and device.type == "cuda" and not t.is_pinned():
            t = t.contiguous().pin_memory()      # (expanded views — M-RoPE position ids — cannot be pinned in place)
        t = t.to(device, non_blocking=non_blocking)
    return t


class _Graph:
    """One captured decode step with its static I/O."""

    def __init__(self):
        self.graph = None
        self.inputs: Dict[str, torch.Tensor] = {}
        self.out: Optional[ModelOutput] = None


class SubModelRunner:
    def __init__(self, tag: str, model, config, batch_size: int, buckets: Sequence, n_active_tokens: int,
                 is_prefill: bool, device: torch.device, forward_kwargs: Optional[dict] = None):
        self.tag = tag
        self.model = model
        self.config = config
        self.neuron_config = nc = config.neuron_config
        self.batch_size = batch_size
        self.buckets = list(buckets)
        self.n_active_tokens = n_active_tokens
        self.is_prefill = is_prefill
        self.device = device
        self.forward_kwargs = forward_kwargs or {}
        self.use_graphs = (device.type == "cuda" and nc.cuda_graphs and not is_prefill
                           and bool(getattr(model, "graph_safe", True)))
        # context encoding is host-launch bound for short prompts (hundreds of eager launches): replay it from a graph too
        self.use_prefill_graphs = (device.type == "cuda" and nc.cuda_graphs and is_prefill and bool(getattr(model, "graph_safe", True))
                                   and nc.torch_dtype == torch.bfloat16 and not nc.is_block_kv_layout and not nc.is_prefix_caching
                                   and os.environ.get("NXDI_B200_PREFILL_GRAPHS", "1") != "0"
                                   and hasattr(model, "_kernels_cover_decode") and model._kernels_cover_decode()
                                   # MoE prefill: only when the routed experts take the grouped-GEMM path (static shapes);
                                   # the PyTorch dispatch has data-dependent shapes
                                   and (not any(getattr(l, "mlp_is_moe", False) for l in getattr(model, "layers", []))
                                        or (hasattr(model, "_moe_kernels_cover_prefill") and model._moe_kernels_cover_prefill())))
        self._prefill_pool = None
        self._graphs: Dict[Tuple, _Graph] = {}
        self.pad_token_id = getattr(config, "pad_token_id", None) or nc.pad_token_id or 0
        self.batch_buckets = sorted(set((nc.token_generation_batches or []) + [batch_size])) \
            if (not is_prefill and nc.token_generation_batches) else [batch_size]
        self.seq_buckets = [b[1] if isinstance(b, (list, tuple)) else b for b in self.buckets]
        self.seq_buckets = sorted(set(self.seq_buckets))
        self.async_feedback = bool(nc.async_mode) and not is_prefill and n_active_tokens == 1
        self.n_launch = 0
        self.collector = None   # utils.benchmark.LatencyCollector while benchmarking
        self.snapshot_hook = None   # utils.snapshot.SnapshotHook
        self.on_new_request = None

    # ------------------------------------------------------------------------------------
    def reset(self):
        pass

    def warmup(self):
        """One forward per bucket (captures graphs).  reference application_base.py:349-373."""
        if self.device.type != "cuda":
            return
        nc = self.neuron_config
        B = self.batch_size
        for sb in self.seq_buckets:
            T = min(sb, nc.max_context_length) if self.is_prefill else self.n_active_tokens
            if self.is_prefill and (nc.is_prefix_caching or nc.is_block_kv_layout):
                continue
            ids = torch.zeros(B, T, dtype=torch.long)
            if self.is_prefill:
                pos = torch.arange(T).unsqueeze(0).expand(B, T).contiguous()
                mask = torch.ones(B, T, dtype=torch.int32)
            else:
                pos = torch.full((B, 1), max(sb - T - 1, 0), dtype=torch.long) + torch.arange(T).unsqueeze(0)
                mask = None
            seq = torch.full((B,), -1, dtype=torch.int32)  # garbage line: do not disturb real caches
            if nc.is_block_kv_layout:
                continue
            self(ids, mask, pos, seq, None, _warmup=True)
        torch.cuda.synchronize()

    # ------------------------------------------------------------------------------------
    def get_target_bucket(self, length: int, strategy: str = "first_fit") -> int:
        nc = self.neuron_config
        if self.is_prefill:
            i = autobucketing.select_prefill_bucket(self.seq_buckets, length, nc.allow_input_truncation)
        else:
            i = autobucketing.select_bucket(self.seq_buckets, length, 0, strategy, nc.allow_input_truncation)
        return self.seq_buckets[i]

    def get_target_2d_bucket_for_prefix_caching(self, active_len: int, prefix_len: int, buckets=None, strategy: str = "first_fit"):
        """(active-token bucket, prefix-length bucket) of a prefix-caching request (reference model_wrapper.py:923-1045).
        Prefill: smallest active bucket holding the new tokens (+ one KV block with EAGLE, whose target recomputes a block); the empty
        tail of that bucket absorbs the end of the cached prefix (recomputed instead of read back), the rest of the prefix picks the
        prefix bucket; 256 < total <= 512 goes to the (512, 0) bucket when it exists.  Decode / speculation: smallest prefix bucket
        strictly longer than the context (+ speculation length) among the rows with enough active slots."""
        from ..modules import autobucketing
        nc = self.neuron_config
        if strategy not in ("first_fit", "second_fit"):
            raise ValueError('Strategy must be either "first_fit" or "second_fit"')
        if buckets is None:
            buckets = autobucketing.generate_2d_buckets_for_prefix_caching(
                min(128, nc.max_context_length), nc.max_context_length, min(128, nc.max_length), nc.max_length, self.is_prefill)
        acts = sorted({b[0] for b in buckets})
        pres = sorted({b[1] for b in buckets})
        if not self.is_prefill:
            spec = 0 if nc.async_mode else (nc.speculation_length if self.n_active_tokens > 1 else 0)
            ok = sorted((p, a) for a, p in buckets if a >= active_len and p > prefix_len + spec)
            if not ok:
                if not nc.allow_input_truncation:
                    raise ValueError(f"context {prefix_len} exceeds the largest bucket ({pres[-1]}) for {self.tag}")
                return [acts[-1], pres[-1]]
            p = ok[0][0]
            if strategy == "second_fit":
                p = next((q for q in pres if q > p), p)
            return [min(a for a, q in buckets if q == ok[0][0] and a >= active_len), p]
        total = active_len + prefix_len
        if 256 < total <= 512 and [512, 0] in [list(b) for b in buckets]:
            return [512, 0]
        blk = nc.pa_block_size if nc.enable_eagle_speculation else 0
        need = active_len + blk
        a = next((x for x in acts if x >= need), None)
        if a is None:
            if not nc.allow_input_truncation:
                raise ValueError(f"prefill length {need} exceeds the largest bucket ({acts[-1]}) for {self.tag}")
            a = acts[-1]
        spare = max(0, a - need)
        if blk:
            spare = (spare // blk) * blk                       # whole blocks only move from the prefix to the prefill side
        rest = max(0, prefix_len - spare)
        p = next((x for x in pres if x >= rest), None)
        if p is None:
            raise ValueError(f"prefix length {rest} exceeds the largest prefix bucket ({pres[-1]}) for {self.tag}")
        return [a, p]

    def vllm_cte_repadding(self, input_ids, attention_mask, position_ids):
        """vLLM pads a batch of prompts to ITS longest prompt; strip that (true length = max position + 1) and pad to this runner's
        nearest bucket instead (reference model_wrapper.py:1297-1313).  -> (ids, mask, positions, true length) like ``pad_prefill``."""
        n = int(position_ids.max()) + 1
        return self.pad_prefill(input_ids[:, :n], attention_mask[:, :n] if attention_mask is not None else None, position_ids[:, :n])

    def pad_prefill(self, input_ids, attention_mask, position_ids):
        """Right/left pad the token axis up to the bucket (reference model_wrapper.py:730-829:
        ids <- pad_token_id, mask <- 0, positions <- 1)."""
        nc = self.neuron_config
        T = input_ids.shape[1]
        tgt = self.get_target_bucket(T)
        if T > tgt:
            if not nc.allow_input_truncation:
                raise ValueError(f"Inputs supplied ({T}) are longer than max_context_length bucket ({tgt})")
            input_ids, position_ids = input_ids[:, :tgt], position_ids[:, :tgt]
            attention_mask = attention_mask[:, :tgt] if attention_mask is not None else None
            return input_ids, attention_mask, position_ids, tgt
        if T == tgt:
            return input_ids, attention_mask, position_ids, T
        pad = tgt - T
        left = nc.padding_side == "left"

        def p(t, v):
            filler = torch.full((t.shape[0], pad), v, dtype=t.dtype, device=t.device)
            return torch.cat([filler, t], 1) if left else torch.cat([t, filler], 1)
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids, dtype=torch.int32)
        return p(input_ids, self.pad_token_id), p(attention_mask, 0), p(position_ids, 1), T

    # ------------------------------------------------------------------------------------
    def __call__(self, *a, **kw):
        col = self.collector
        if col is None:
            return self._call(*a, **kw)
        col.pre_hook()
        out = self._call(*a, **kw)
        col.hook()
        return out

    def _call(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params, _warmup=False, **kw):
        B = input_ids.shape[0]
        if self.snapshot_hook is not None and not _warmup:
            if self.is_prefill and self.on_new_request is not None:
                self.on_new_request()
            h = self.snapshot_hook
            if self.is_prefill:
                h.request -= 1          # on_new_request advanced every runner; the hook itself advances on prefill
            h(dict(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids, seq_ids=seq_ids,
                   sampling_params=sampling_params), self.is_prefill)
        if B > self.batch_size:
            # request larger than the compiled batch: run chunks sequentially (model_wrapper.py:1358-1423)
            outs = []
            for s in range(0, B, self.batch_size):
                sl = slice(s, s + self.batch_size)
                sub = {k: (v[sl] if torch.is_tensor(v) and v.shape[:1] == (B,) else v) for k, v in kw.items()}
                o = self._call(input_ids[sl], None if attention_mask is None else attention_mask[sl],
                               position_ids[sl], seq_ids[sl],
                               None if sampling_params is None else sampling_params[sl], **sub)
                # graph-backed outputs are views of the graph's static buffers: the next chunk's replay overwrites them
                outs.append(_clone_output(o))
            return _cat_outputs(outs)
        if self.is_prefill:
            return self._run_prefill(input_ids, attention_mask, position_ids, seq_ids, sampling_params, **kw)
        return self._run_decode(input_ids, attention_mask, position_ids, seq_ids, sampling_params, **kw)

    def _common_kwargs(self, kw):
        dev = self.device
        out = dict(self.forward_kwargs)
        for k, v in kw.items():
            if v is None:
                continue
            out[k] = _to_dev(v, dev) if torch.is_tensor(v) else v
        return out

    def _run_prefill(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params, **kw):
        dev = self.device
        if attention_mask is not None and attention_mask.shape[-1] != input_ids.shape[-1]:
            attention_mask = attention_mask[:, -input_ids.shape[-1]:]
        T0 = input_ids.shape[1]
        ids, mask, pos, _ = self.pad_prefill(input_ids, attention_mask, position_ids)
        sm = kw.get("slot_mapping")
        if sm is not None and ids.shape[1] != T0:
            # padded tokens must not be written: slot -1 (reference pads the slot mapping the same way)
            fill = torch.full((sm.shape[0], ids.shape[1] - T0), -1, dtype=sm.dtype, device=sm.device)
            left = self.neuron_config.padding_side == "left"
            kw["slot_mapping"] = torch.cat([fill, sm], 1) if left else torch.cat([sm, fill], 1)
        if ids.shape[1] != T0:
            padn = ids.shape[1] - T0
            left = self.neuron_config.padding_side == "left"
            for name, fill in (("vision_mask", 0), ("rotary_position_ids", 1)):
                t = kw.get(name)
                if t is not None and t.shape[-1] == T0:
                    f = torch.full(tuple(t.shape[:-1]) + (padn,), fill, dtype=t.dtype, device=t.device)
                    kw[name] = torch.cat([f, t], -1) if left else torch.cat([t, f], -1)
        if mask is not None and mask.device.type == "cpu" and bool(mask.all()):
            mask = None   # no padding anywhere: skip the mask plumbing (decided on the host, no device sync)
        self.n_launch += 1
        kwargs = self._common_kwargs(kw)
        if self.use_prefill_graphs and not any(torch.is_tensor(v) for v in kwargs.values()) and not kwargs.get("has_prefix") \
                and ids.shape[1] <= 2048:
            try:
                return self._run_prefill_graph(ids, mask, pos, seq_ids, sampling_params, kwargs)
            except RuntimeError as e:
                # an op of this configuration is not capturable (same policy as the decode runner): eager context encoding from now on
                logger.warning("%s: CUDA-graph capture of context encoding failed (%s); falling back to eager prefill", self.tag,
                               str(e).split("\n")[0])
                torch.cuda.synchronize(dev)
                self.use_prefill_graphs = False
        with torch.no_grad():
            return self.model(_to_dev(ids, dev), _to_dev(mask, dev), _to_dev(pos, dev, torch.int32),
                              _to_dev(seq_ids, dev, torch.int32), _to_dev(sampling_params, dev, torch.float32),
                              is_prefill=True, **kwargs)

    def _run_prefill_graph(self, ids, mask, pos, seq_ids, sampling_params, kwargs):
        dev = self.device
        B, T = ids.shape
        key = ("cte", B, T, mask is not None, tuple(sorted((k, v) for k, v in kwargs.items())))
        g = self._graphs.get(key)
        if g is None:
            g = _Graph()
            g.inputs = dict(input_ids=torch.zeros(B, T, dtype=torch.long, device=dev),
                            position_ids=torch.zeros(B, T, dtype=torch.int32, device=dev),
                            seq_ids=torch.full((B,), -1, dtype=torch.int32, device=dev),
                            sampling_params=torch.tensor([[1.0, 1.0, 1.0]], device=dev).repeat(B, 1))
            if mask is not None:
                g.inputs["mask"] = torch.ones(B, T, dtype=torch.int32, device=dev)
            si = g.inputs
            si["position_ids"].copy_(torch.arange(T, device=dev).view(1, T).expand(B, T))

            def run():
                return self.model(si["input_ids"], si.get("mask"), si["position_ids"], si["seq_ids"], si["sampling_params"],
                                  is_prefill=True, **kwargs)
            st = torch.cuda.Stream(device=dev)
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.no_grad(), torch.cuda.stream(st):
                run()          # warm-up outside capture (lazy tables, workspaces); writes go to the garbage line
            torch.cuda.current_stream(dev).wait_stream(st)
            torch.cuda.synchronize(dev)
            self._symm_even()
            graph = torch.cuda.CUDAGraph()
            if self._prefill_pool is None:
                self._prefill_pool = torch.cuda.graph_pool_handle()
            with torch.no_grad(), torch.cuda.graph(graph, stream=st, pool=self._prefill_pool):
                g.out = run()
                self._symm_even()
            g.graph = graph
            self._graphs[key] = g
            logger.debug("captured %s prefill graph B=%d T=%d", self.tag, B, T)
        si = g.inputs
        si["input_ids"].copy_(ids, non_blocking=True)
        si["position_ids"].copy_(pos, non_blocking=True)
        si["seq_ids"].copy_(seq_ids, non_blocking=True)
        if mask is not None:
            si["mask"].copy_(mask, non_blocking=True)
        if sampling_params is not None:
            si["sampling_params"].copy_(sampling_params, non_blocking=True)
        self._symm_even()
        g.graph.replay()
        o = g.out
        # logits / hidden states are handed out as copies: the graph's static output buffer is overwritten by the next replay, and a
        # caller that keeps `out.logits` across calls (accuracy checks, logit matching) must not see it change under its feet.  The
        # sampled tokens stay a view (8 bytes per row, consumed immediately by every loop in this repo; the async path reads them in place).
        return ModelOutput(tokens=o.tokens, logits=None if o.logits is None else o.logits.clone(),
                           hidden_states=None if o.hidden_states is None else o.hidden_states.clone())

    # ---- decode ------------------------------------------------------------------------------
    def _pick_batch_bucket(self, B):
        for b in self.batch_buckets:
            if B <= b:
                return b
        return self.batch_buckets[-1]

    def _run_decode(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params, **kw):
        dev = self.device
        B, T = input_ids.shape
        kwargs = self._common_kwargs(kw)
        graphable = self.use_graphs and not any(torch.is_tensor(v) for v in kwargs.values())
        if not graphable:
            self.n_launch += 1
            with torch.no_grad():
                return self.model(_to_dev(input_ids, dev), None, _to_dev(position_ids, dev, torch.int32),
                                  _to_dev(seq_ids, dev, torch.int32), _to_dev(sampling_params, dev, torch.float32),
                                  is_prefill=False, **kwargs)
        Bb = self._pick_batch_bucket(B)
        if position_ids.device.type == "cpu":
            cur = max(map(max, position_ids.tolist())) + 1   # plain Python: torch reductions on pinned tensors cost ms
        else:
            cur = self.seq_buckets[-1] - 1  # device-resident inputs: no sync, take the largest bucket
        sb = self.get_target_bucket(cur)
        key = (Bb, sb, T, tuple(sorted((k, v) for k, v in kwargs.items() if not torch.is_tensor(v))))
        g = self._graphs.get(key)
        if g is None:
            try:
                g = self._capture(key, Bb, T, kwargs)
            except RuntimeError as e:
                # an op of this configuration's fallback path is not capturable: run this runner eagerly from now on
                logger.warning("%s: CUDA-graph capture failed (%s); falling back to eager decode", self.tag, str(e).split("\n")[0])
                torch.cuda.synchronize(dev)
                self.use_graphs = False
                return self._run_decode(input_ids, attention_mask, position_ids, seq_ids, sampling_params, **kw)
        # stage inputs into the static buffers
        si = g.inputs
        si["input_ids"][:B].copy_(input_ids, non_blocking=True)
        si["position_ids"][:B].copy_(position_ids, non_blocking=True)
        si["seq_ids"][:B].copy_(seq_ids, non_blocking=True)
        if B < Bb:
            si["seq_ids"][B:].fill_(-1)
            si["position_ids"][B:].fill_(0)
        if sampling_params is not None:
            si["sampling_params"][:B].copy_(sampling_params, non_blocking=True)
        self._symm_even()
        g.graph.replay()
        self.n_launch += 1
        o = g.out
        return ModelOutput(tokens=None if o.tokens is None else o.tokens[:B],
                           logits=None if o.logits is None else o.logits[:B].clone(),
                           hidden_states=None if o.hidden_states is None else o.hidden_states[:B].clone())

    def _capture(self, key, Bb, T, kwargs) -> _Graph:
        dev = self.device
        g = _Graph()
        g.inputs = dict(
            input_ids=torch.zeros(Bb, T, dtype=torch.long, device=dev),
            position_ids=torch.zeros(Bb, T, dtype=torch.int32, device=dev),
            seq_ids=torch.full((Bb,), -1, dtype=torch.int32, device=dev),
            sampling_params=torch.tensor([[1.0, 1.0, 1.0]], device=dev).repeat(Bb, 1),
        )
        si = g.inputs

        def step():
            # seq_hint: the graph of sequence bucket `sb` bakes a split-KV grid sized for sb keys, not for the whole cache
            out = self.model(si["input_ids"], None, si["position_ids"], si["seq_ids"], si["sampling_params"],
                             is_prefill=False, seq_hint=int(key[1]), **kwargs)
            if self.async_feedback and out.tokens is not None:
                si["input_ids"].copy_(out.tokens.view(Bb, 1))
                si["position_ids"].add_(1)
            return out
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(s):
            for _ in range(2):  # warm up (lazy tables, workspace allocs) outside capture
                step()
            si["position_ids"].zero_()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self._symm_even()
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph, stream=s):
            g.out = step()
            self._symm_even()   # captured: keeps the number of fused collectives per replay even
        g.graph = graph
        self._graphs[key] = g
        logger.debug("captured %s graph %s", self.tag, key[:3])
        return g

    def _symm_even(self):
        from ..parallel.state import get_tensor_model_parallel_group
        symm = get_tensor_model_parallel_group().symm
        if symm is not None:
            symm.ensure_even()

    # ---- device-resident multi-step decode (async mode / benchmarks) ------------------------------
    def replay_steps(self, key_graph: _Graph, n: int):
        for _ in range(n):
            key_graph.graph.replay()
        self.n_launch += n

    def graph_for(self, B: int, T: int = 1, cur_len: Optional[int] = None, **kwargs) -> _Graph:
        Bb = self._pick_batch_bucket(B)
        sb = self.get_target_bucket(cur_len if cur_len is not None else self.seq_buckets[-1] - 1)
        key = (Bb, sb, T, tuple(sorted(kwargs.items())))
        g = self._graphs.get(key)
        if g is None:
            g = self._capture(key, Bb, T, kwargs)
        return g


def _clone_output(o: ModelOutput) -> ModelOutput:
    c = lambda t: None if t is None else t.clone()
    return ModelOutput(tokens=c(o.tokens), logits=c(o.logits), hidden_states=c(o.hidden_states))


def _cat_outputs(outs: List[ModelOutput]) -> ModelOutput:
    def cat(name):
        vals = [getattr(o, name) for o in outs]
        return None if vals[0] is None else torch.cat(vals, 0)
    return ModelOutput(tokens=cat("tokens"), logits=cat("logits"), hidden_states=cat("hidden_states"))
