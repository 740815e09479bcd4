"""Speculative decoding: vanilla (separate draft application), fused (draft + target in one device step with
on-device acceptance) and the host loops that drive them.

reference: ``NeuronFusedSpecModel`` (models/model_base.py:1598-3021: ``_token_gen_forward`` :1812-1929,
``_tkg_postprocessor`` :2799-2855) and the HF-adapter loops ``_standard_assisted_decoding`` /
``_fused_assisted_decoding`` (utils/hf_adapter.py:495-797).

Conventions (kept from the reference): ``speculation_length = k``; per step the draft proposes ``k-1`` tokens from
the last accepted token, the target scores the ``k`` tokens ``[last, d1..d_{k-1}]`` in ONE forward with
``n_active_tokens = k``, the longest matching prefix is accepted plus one bonus token from the target, and the
step returns ``accepted_tokens`` padded with ``-1`` together with the next inputs — all computed on the device.
KV caches are never rolled back: rejected positions are simply overwritten by the next step, because every
kernel addresses the cache by absolute position.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn


@torch.no_grad()
def greedy_accept(draft_tokens: torch.Tensor, target_tokens: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """draft_tokens [B,k-1], target_tokens [B,k] (target's arg-max after each of the k inputs).
    -> (accepted [B,k] padded with -1, n_accepted [B] in 1..k).  Pure tensor ops (graph capturable)."""
    B, k = target_tokens.shape
    if k == 1:
        return target_tokens.clone(), torch.ones(B, dtype=torch.long, device=target_tokens.device)
    match = (draft_tokens == target_tokens[:, : k - 1]).long()
    n_match = match.cumprod(-1).sum(-1)                      # leading matches, 0..k-1
    n_acc = n_match + 1                                       # + bonus token
    ar = torch.arange(k, device=target_tokens.device).view(1, k)
    accepted = torch.where(ar < n_acc.view(B, 1), target_tokens, torch.full_like(target_tokens, -1))
    return accepted, n_acc


@torch.no_grad()
def adjust_target_probs(target_probs: torch.Tensor, draft_probs: torch.Tensor) -> torch.Tensor:
    """Residual distribution ``norm(max(0, p_target - p_draft))`` sampled after a rejection (reference
    ``_adjust_target_probs`` model_base.py:1678-1695)."""
    r = (target_probs - draft_probs).clamp_min(0)
    z = r.sum(-1, keepdim=True)
    return torch.where(z > 0, r / z.clamp_min(1e-30), target_probs)


@torch.no_grad()
def speculative_sample_accept(draft_tokens: torch.Tensor, draft_probs: torch.Tensor, target_probs: torch.Tensor,
                              rand_accept: torch.Tensor, rand_sample: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stochastic (rejection-sampling) acceptance — the output follows the TARGET distribution exactly.
    draft_tokens [B,k-1]; draft_probs [B,k-1,V] (distribution each draft token was drawn from); target_probs [B,k,V] (target
    distribution after each of the k inputs); rand_accept [B,k-1], rand_sample [B,k] uniforms.
    Token i is kept iff ``u_i < p_t(x_i) / p_d(x_i)``; at the first rejection a replacement is drawn from the residual
    distribution, after k-1 acceptances a bonus token is drawn from the last target distribution.
    -> (accepted [B,k] padded with -1, n_accepted [B] in 1..k).  Static shapes, no host sync."""
    B, km1, V = draft_probs.shape
    k = km1 + 1
    dev = draft_tokens.device
    idx = draft_tokens.unsqueeze(-1)
    pt = target_probs[:, :km1].gather(-1, idx).squeeze(-1)
    pd = draft_probs.gather(-1, idx).squeeze(-1).clamp_min(1e-30)
    ok = rand_accept < (pt / pd)
    n_match = ok.long().cumprod(-1).sum(-1)                                   # 0..k-1 draft tokens kept
    # distribution of the token that follows the kept prefix
    resid = adjust_target_probs(target_probs[:, :km1], draft_probs)           # used when a rejection happened at that slot
    nxt = torch.cat([resid, target_probs[:, km1:]], 1)                        # [B,k,V]; slot k-1 = bonus (no rejection)
    dist = nxt.gather(1, n_match.view(B, 1, 1).expand(B, 1, V)).squeeze(1)
    u = rand_sample.gather(1, n_match.view(B, 1))
    cdf = dist.cumsum(-1)
    new_tok = (cdf < u * cdf[:, -1:]).sum(-1).clamp_max(V - 1)
    ar = torch.arange(k, device=dev).view(1, k)
    toks = torch.cat([draft_tokens, draft_tokens.new_zeros(B, 1)], 1)
    accepted = torch.where(ar < n_match.view(B, 1), toks, torch.full_like(toks, -1))
    accepted = torch.where(ar == n_match.view(B, 1), new_tok.view(B, 1), accepted)
    return accepted, n_match + 1


@torch.no_grad()
def filtered_probs(logits: torch.Tensor, sampling_params: torch.Tensor) -> torch.Tensor:
    """Per-request sampling distribution: temperature, top-k, top-p (the on-device sampler's rules, modules/sampling.py) as a
    dense probability vector.  logits [B,T,V], sampling_params [B,3] = (top_k, top_p, temperature)."""
    B, T, V = logits.shape
    top_k = sampling_params[:, 0].view(B, 1, 1)
    top_p = sampling_params[:, 1].view(B, 1, 1)
    temp = sampling_params[:, 2].view(B, 1, 1)
    greedy = (temp == 0) | (top_k == 1)
    x = logits.float() / torch.where(temp > 0, temp, torch.ones_like(temp))
    sv, si = x.sort(-1, descending=True)
    rank = torch.arange(V, device=logits.device).view(1, 1, V)
    keep = torch.where(top_k > 0, rank < top_k, torch.ones_like(rank, dtype=torch.bool)).expand(B, T, V)
    p = torch.softmax(sv.masked_fill(~keep, float("-inf")), -1)
    keep = (keep & ((p.cumsum(-1) - p) < top_p)) | (rank == 0)
    keep = torch.where(greedy, rank == 0, keep)
    p = torch.softmax(sv.masked_fill(~keep, float("-inf")), -1)
    return torch.zeros_like(p).scatter(-1, si, p)


class FusedSpeculativeModel(nn.Module):
    """Draft + target in one step (role of ``NeuronFusedSpecModel``).  ``forward`` runs k-1 greedy draft steps, one
    target verify over k tokens, the acceptance rule and the next-input derivation without any host
    synchronisation, so the whole step is a single CUDA-graph replay."""

    def __init__(self, target: nn.Module, draft: nn.Module, speculation_length: int):
        super().__init__()
        self.target_model = target
        self.draft_model = draft
        self.k = speculation_length

    def reset(self):
        self.target_model.reset()
        self.draft_model.reset()

    @torch.no_grad()
    def prefill(self, input_ids, attention_mask, position_ids, seq_ids, sampling_params=None, **kw):
        """Context encoding of both models on the same prompt (reference :1750-1810); returns the target's token."""
        out_t = self.target_model(input_ids, attention_mask, position_ids, seq_ids, sampling_params, is_prefill=True, **kw)
        self.draft_model(input_ids, attention_mask, position_ids, seq_ids, None, is_prefill=True)
        return out_t

    @torch.no_grad()
    def forward_sampling(self, last_token, position, seq_ids, prev_token, sampling_params, generator=None):
        """Stochastic variant (``do_sample``): draft tokens are SAMPLED from the filtered draft distribution, the target verifies
        with rejection sampling (``speculative_sample_accept``) — the emitted tokens follow the target's sampling distribution
        (reference ``_token_gen_forward`` with ``_adjust_target_probs``, model_base.py:1678-1695,1812-1929)."""
        k, B, dev = self.k, last_token.shape[0], last_token.device
        sp = sampling_params.to(dev).float()
        u = lambda *shape: torch.rand(*shape, device=dev, generator=generator)   # noqa: E731

        def draw(probs, r):
            cdf = probs.cumsum(-1)
            return (cdf < r.unsqueeze(-1) * cdf[..., -1:]).sum(-1).clamp_max(probs.shape[-1] - 1)
        ids = torch.cat([prev_token, last_token], 1)
        out = self.draft_model(ids, None, torch.cat([position - 1, position], 1), seq_ids, None, is_prefill=False,
                               all_positions=True, output_logits=True)
        dprobs, dtoks = [], []
        pd = filtered_probs(out.logits[:, -1:], sp)
        pos = position + 1
        for i in range(k - 1):
            tok = draw(pd, u(B, 1))
            dprobs.append(pd)
            dtoks.append(tok)
            if i == k - 2:
                break
            out = self.draft_model(tok, None, pos, seq_ids, None, is_prefill=False, output_logits=True)
            pd = filtered_probs(out.logits[:, -1:], sp)
            pos = pos + 1
        draft_tokens = torch.cat(dtoks, 1)
        cand_ids = torch.cat([last_token, draft_tokens], 1)
        cand_pos = position + torch.arange(k, device=dev, dtype=position.dtype).view(1, k)
        out_t = self.target_model(cand_ids, None, cand_pos, seq_ids, None, is_prefill=False, all_positions=True, output_logits=True)
        pt = filtered_probs(out_t.logits, sp)
        accepted, n_acc = speculative_sample_accept(draft_tokens, torch.cat(dprobs, 1), pt, u(B, k - 1), u(B, k))
        next_token = accepted.gather(1, (n_acc - 1).view(B, 1))
        next_pos = position + n_acc.view(B, 1).to(position.dtype)
        prev_in_step = accepted.gather(1, (n_acc - 2).clamp_min(0).view(B, 1))
        next_prev = torch.where(n_acc.view(B, 1) >= 2, prev_in_step, last_token)
        return accepted, n_acc, next_token, next_pos, next_prev, out_t

    @torch.no_grad()
    def forward(self, last_token: torch.Tensor, position: torch.Tensor, seq_ids: torch.Tensor,
                prev_token: torch.Tensor, sampling_params: Optional[torch.Tensor] = None, generator=None):
        if sampling_params is not None and bool(getattr(self.target_model.neuron_config.on_device_sampling_config, "do_sample", False)):
            return self.forward_sampling(last_token, position, seq_ids, prev_token, sampling_params, generator)
        """last_token [B,1] at absolute ``position`` [B,1]; ``prev_token`` [B,1] is the token at ``position-1``.
        The first draft step always feeds ``[prev_token, last_token]``: when the previous step accepted every
        draft token the draft has not seen ``prev_token`` yet; otherwise it rewrites an identical KV entry, so no
        host-side branching is needed and the step is shape-static."""
        k = self.k
        B = last_token.shape[0]
        dev = last_token.device
        cand = [last_token]
        ids = torch.cat([prev_token, last_token], 1)
        p2 = torch.cat([position - 1, position], 1)
        out = self.draft_model(ids, None, p2, seq_ids, None, is_prefill=False, all_positions=True)
        tok = _last_tokens(out)[:, -1:]
        pos = position + 1
        for i in range(k - 1):
            cand.append(tok)
            if i == k - 2:
                break
            out = self.draft_model(tok, None, pos, seq_ids, None, is_prefill=False)
            tok = _last_tokens(out)[:, -1:]
            pos = pos + 1
        cand_ids = torch.cat(cand, 1)                                          # [B,k]
        cand_pos = position + torch.arange(k, device=dev, dtype=position.dtype).view(1, k)
        # ---- one target verify over the k tokens ----
        out_t = self.target_model(cand_ids, None, cand_pos, seq_ids, None, is_prefill=False, all_positions=True)
        tgt = _last_tokens(out_t)                                              # [B,k]
        accepted, n_acc = greedy_accept(cand_ids[:, 1:], tgt)
        # ---- next inputs on device (reference _tkg_postprocessor :2799-2855) ----
        next_token = tgt.gather(1, (n_acc - 1).view(B, 1))
        next_pos = position + n_acc.view(B, 1).to(position.dtype)
        prev_in_step = tgt.gather(1, (n_acc - 2).clamp_min(0).view(B, 1))
        next_prev = torch.where(n_acc.view(B, 1) >= 2, prev_in_step, last_token)
        return accepted, n_acc, next_token, next_pos, next_prev, out_t


def _last_tokens(out) -> torch.Tensor:
    if out.tokens is not None:
        t = out.tokens
        return t.view(t.shape[0], -1)
    return out.logits.argmax(-1)


class GraphedStep:
    """One fused speculation step captured as a CUDA graph: static input buffers, one ``replay()`` per step.

    The reference compiles draft + verify + acceptance into ONE device graph (``NeuronFusedSpecModel``); here the step functions
    are already free of host synchronisation (acceptance, KV compaction and rolling-buffer updates are tensor code), so the
    eager Python that issues their ~hundreds of launches is replaced by a graph replay.  Warm-up and capture run with
    ``seq_ids = -1`` so that every cache / buffer write lands on the garbage line."""

    def __init__(self, fn, example_inputs, n_outputs: int, seq_ids_index: int = 2):
        self.fn, self.n_out = fn, n_outputs
        dev = example_inputs[0].device
        self.static_in = [t.clone() for t in example_inputs]
        warm = [t.clone() for t in example_inputs]
        warm[seq_ids_index] = torch.full_like(warm[seq_ids_index], -1)
        st = torch.cuda.Stream(device=dev)
        st.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(st):
            for _ in range(2):
                fn(*warm)
        torch.cuda.current_stream(dev).wait_stream(st)
        torch.cuda.synchronize(dev)
        for dst, src in zip(self.static_in, warm):
            dst.copy_(src)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph, stream=st):
            out = fn(*self.static_in)
        self.static_out = tuple(out[: n_outputs])

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


def maybe_graph_step(owner, fn, example_inputs, n_outputs: int, key):
    """Graph ``fn`` when the device / model allow it (cached on ``owner`` per ``key``); otherwise return ``fn`` itself."""
    import os
    dev = example_inputs[0].device
    tm = getattr(owner, "target_model", None)
    dm = getattr(owner, "draft_model", None)      # a draft on the PyTorch composite path (odd head_dim, ...) is not capturable either
    if (dev.type != "cuda" or os.environ.get("NXDI_B200_SPEC_GRAPH", "1") == "0" or tm is None or not getattr(tm, "graph_safe", False)
            or not tm.neuron_config.cuda_graphs or (dm is not None and not getattr(dm, "graph_safe", True))):
        return fn
    cache = owner.__dict__.setdefault("_graphed_steps", {})
    if key not in cache:
        cache[key] = GraphedStep(fn, example_inputs, n_outputs)
    return cache[key]


# ---- host loops ----------------------------------------------------------------------------------------------------
@torch.no_grad()
def assisted_generate(adapter, input_ids, attention_mask, max_length, eos: List[int], pad_id: int, assistant_model=None,
                      sampling_params=None, gc=None, return_dict_in_generate: bool = False):
    """Greedy speculative generation.  ``assistant_model``: a separate draft application (vanilla speculation);
    when the target was built with ``enable_fused_speculation`` its own fused model is used instead."""
    model = adapter.neuron_model
    nc = model.neuron_config
    k = max(nc.speculation_length, 2)
    B, P = input_ids.shape
    fused = getattr(model, "fused_spec_model", None)
    if fused is None:
        if assistant_model is None:
            raise ValueError("speculation needs a draft: pass assistant_model or enable_fused_speculation")
        fused = FusedSpeculativeModel(model.model, assistant_model.model, k)
    dev = model.device
    model.reset()
    if assistant_model is not None:
        assistant_model.reset()
    seq_ids = torch.arange(B, dtype=torch.int32, device=dev)
    pos0 = (attention_mask.long().cumsum(-1) - 1).clamp_min(0).to(torch.int32)
    out = fused.prefill(input_ids.to(dev), attention_mask.to(dev), pos0.to(dev), seq_ids)
    tok = _last_tokens(out)[:, -1:].to(dev)
    n_valid = attention_mask.sum(-1).view(B, 1).to(dev)
    position = n_valid.to(torch.int32)           # position of `tok`
    prev = input_ids.to(dev).gather(1, (n_valid.long() - 1).clamp_min(0))
    rows = [input_ids[b, : int(n_valid[b])].tolist() + [int(tok[b])] for b in range(B)]
    done = [int(tok[b]) in eos for b in range(B)]
    stats = {"steps": 0, "accepted": 0}
    sampled = (sampling_params is not None and isinstance(fused, FusedSpeculativeModel)
               and bool(getattr(nc.on_device_sampling_config, "do_sample", False)))
    step = fused if sampled else maybe_graph_step(fused, lambda t, p, s, pv: fused(t, p, s, pv), [tok, position, seq_ids, prev], 5,
                                                  ("greedy", B))
    while not all(done) and min(len(r) for r, d in zip(rows, done) if not d) < max_length:
        if sampled:
            accepted, n_acc, tok, position, prev, _ = fused(tok, position, seq_ids, prev, sampling_params)
        else:
            accepted, n_acc, tok, position, prev = step(tok, position, seq_ids, prev)[:5]
        acc = accepted.cpu()
        stats["steps"] += 1
        stats["accepted"] += int(n_acc.sum())
        for b in range(B):
            if done[b]:
                continue
            for t in acc[b].tolist():
                if t < 0:
                    break
                if len(rows[b]) >= max_length:
                    done[b] = True
                    break
                rows[b].append(t)
                if t in eos:
                    done[b] = True
                    break
            if len(rows[b]) >= max_length:
                done[b] = True
    width = max(len(r) for r in rows)
    seqs = torch.full((B, width), pad_id, dtype=torch.long)
    for b, r in enumerate(rows):
        seqs[b, : len(r)] = torch.tensor(r)
    if return_dict_in_generate:
        from ..utils.hf_adapter import GenerateOutput
        return GenerateOutput(sequences=seqs, speculation_stats=stats)
    return seqs
