"""Whisper decoding task (reference models/whisper/utils/decoding.py plugs the Neuron model into openai-whisper's ``DecodingTask``;
openai-whisper is not a dependency here, so the decoding rules themselves are implemented):

* logit filters — suppressed tokens, tokens suppressed at the first position, and the timestamp grammar (timestamps come in pairs,
  never decrease, the first generated token is a timestamp no later than ``max_initial_timestamp_index``, and a timestamp is forced
  when the timestamp tokens together are more likely than any text token);
* greedy or temperature sampling with per-sequence sum / average log-probability;
* the temperature-fallback loop: a segment is re-decoded at the next temperature when its average log-probability is below
  ``logprob_threshold`` or its token sequence compresses too well (repetition loops)."""
from __future__ import annotations

import zlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch


@dataclass
class DecodingOptions:
    max_new_tokens: int = 224
    temperatures: Tuple[float, ...] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0)
    suppress_tokens: Sequence[int] = ()
    begin_suppress_tokens: Sequence[int] = ()
    without_timestamps: bool = True
    no_timestamps_token_id: Optional[int] = None
    max_initial_timestamp_index: Optional[int] = 50
    eos_token_id: Optional[int] = None
    logprob_threshold: Optional[float] = -1.0
    compression_ratio_threshold: Optional[float] = 2.4
    seed: int = 0


@dataclass
class DecodingResult:
    tokens: List[int] = field(default_factory=list)
    sum_logprob: float = 0.0
    avg_logprob: float = 0.0
    temperature: float = 0.0
    compression_ratio: float = 1.0


def suppress_tokens(scores: torch.Tensor, ids: Sequence[int]) -> torch.Tensor:
    if len(ids):
        scores[:, torch.as_tensor(list(ids), device=scores.device)] = float("-inf")
    return scores


def apply_timestamp_rules(scores: torch.Tensor, generated: torch.Tensor, no_timestamps_token_id: int, eos_token_id: int,
                          max_initial_timestamp_index: Optional[int] = None) -> torch.Tensor:
    """``generated`` [B, n]: tokens produced so far (prompt excluded).  Timestamp tokens are the ids above ``no_timestamps_token_id``."""
    ts0 = no_timestamps_token_id + 1
    s = scores.clone()
    s[:, no_timestamps_token_id] = float("-inf")
    B, n = generated.shape
    V = s.shape[1]
    col = torch.arange(V, device=s.device).view(1, V)
    if n >= 1:
        is_ts = generated >= ts0
        last = is_ts[:, -1]
        penult = is_ts[:, -2] if n >= 2 else torch.ones(B, dtype=torch.bool, device=s.device)
        s = s.masked_fill((last & penult).view(B, 1) & (col >= ts0), float("-inf"))              # after a pair: text (or eos) only
        s = s.masked_fill((last & ~penult).view(B, 1) & (col < eos_token_id), float("-inf"))      # an opened pair must be closed
        has = is_ts.any(1)
        last_ts = torch.where(is_ts, generated, torch.zeros_like(generated)).max(1).values      # timestamps never decrease
        floor = torch.where(last & ~penult, last_ts, last_ts + 1)
        s = s.masked_fill(has.view(B, 1) & (col >= ts0) & (col < floor.view(B, 1)), float("-inf"))
    else:
        s[:, :ts0] = float("-inf")                                                               # the first token is a timestamp
        if max_initial_timestamp_index is not None:
            s[:, ts0 + max_initial_timestamp_index + 1:] = float("-inf")
    lp = torch.log_softmax(s.float(), -1)
    force = lp[:, ts0:].logsumexp(-1) > lp[:, :ts0].max(-1).values
    return s.masked_fill(force.view(B, 1) & (col < ts0), float("-inf"))


def compression_ratio(tokens: Sequence[int]) -> float:
    raw = b"".join(int(t).to_bytes(4, "little") for t in tokens)
    return len(raw) / max(len(zlib.compress(raw)), 1) if raw else 1.0


@torch.no_grad()
def _decode_once(app, input_features, prompt, opt: DecodingOptions, temperature: float, gen: torch.Generator):
    B = input_features.shape[0]
    eos = opt.eos_token_id if opt.eos_token_id is not None else getattr(app.config, "eos_token_id")
    app.reset()
    out = app(prompt, input_features=input_features, output_logits=True)
    generated = torch.zeros(B, 0, dtype=torch.long)
    sum_lp = torch.zeros(B)
    done = torch.zeros(B, dtype=torch.bool)
    for step in range(opt.max_new_tokens):
        scores = out.logits[:, -1].float().cpu().clone()
        suppress_tokens(scores, opt.suppress_tokens)
        if step == 0:
            suppress_tokens(scores, opt.begin_suppress_tokens)
        if not opt.without_timestamps:
            scores = apply_timestamp_rules(scores, generated, opt.no_timestamps_token_id, eos, opt.max_initial_timestamp_index)
        lp = torch.log_softmax(scores, -1)
        nxt = lp.argmax(-1) if temperature == 0 else torch.multinomial(torch.softmax(scores / temperature, -1), 1, generator=gen).view(-1)
        nxt = torch.where(done, torch.full_like(nxt, eos), nxt)
        sum_lp += torch.where(done, torch.zeros(B), lp.gather(1, nxt.view(B, 1)).view(B))
        generated = torch.cat([generated, nxt.view(B, 1)], 1)
        done |= nxt == eos
        if bool(done.all()):
            break
        pos = torch.full((B, 1), prompt.shape[1] + step, dtype=torch.int32)
        out = app(nxt.view(B, 1), position_ids=pos, output_logits=True)
    res = []
    for b in range(B):
        toks = generated[b].tolist()
        if eos in toks:
            toks = toks[: toks.index(eos)]
        res.append(DecodingResult(toks, float(sum_lp[b]), float(sum_lp[b]) / (len(toks) + 1), temperature, compression_ratio(toks)))
    return res


def decode(app, input_features: torch.Tensor, decoder_input_ids: torch.Tensor, options: Optional[DecodingOptions] = None) -> List[DecodingResult]:
    """Decode a batch of 30-second segments with temperature fallback (each segment keeps its first acceptable result)."""
    opt = options or DecodingOptions()
    gen = torch.Generator().manual_seed(opt.seed)
    B = input_features.shape[0]
    final: List[Optional[DecodingResult]] = [None] * B
    todo = list(range(B))
    for t in opt.temperatures:
        res = _decode_once(app, input_features[todo], decoder_input_ids[todo], opt, t, gen)
        again = []
        for i, r in zip(todo, res):
            bad = ((opt.compression_ratio_threshold is not None and r.compression_ratio > opt.compression_ratio_threshold)
                   or (opt.logprob_threshold is not None and r.avg_logprob < opt.logprob_threshold))
            final[i] = r
            if bad:
                again.append(i)
        todo = again
        if not todo:
            break
    return final
