"""Accuracy harness: token matching, teacher-forced logit matching with top-k tolerance tiers, draft-logit
matching.  Role of reference utils/accuracy.py:244-1277 and of the in-repo validation algorithm
experimental/core/accuracy/logit_validation.py:73-346 (defaults :14-21).

Logit matching semantics kept from the reference:
  * goldens are the HF-CPU (fp32) greedy logits ``[steps, B, V]``;
  * the model under test is *teacher forced*: whenever its arg-max would leave the golden sequence, the
    golden token is fed instead, so every position is compared under identical context;
  * per position the error is checked on nested top-k slices of the golden distribution with tolerances that
    tighten towards the head: ``{"5": (1e-5, .01), "50": (1e-5, .02), "1000": (1e-5, .03), "all": (1e-5, .05)}``
    after removing the mean shift between the two logit vectors (softmax is shift invariant);
  * at a divergence (different arg-max) the run still passes when the golden token's logit is within
    ``divergence_difference_tol`` of the model's own best logit (a near-tie, not an error).
"""
from __future__ import annotations

import logging
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch

from .hf_adapter import HuggingFaceGenerationAdapter

logger = logging.getLogger("b200infer")

DEFAULT_TOLERANCE_MAP = {"5": (1e-5, 0.01), "50": (1e-5, 0.02), "1000": (1e-5, 0.03), "all": (1e-5, 0.05)}
DEFAULT_DIVERGENCE_DIFFERENCE_TOLERANCE = 0.001


class LogitMatchingValidationError(AssertionError):
    def __init__(self, message, results=None):
        super().__init__(message)
        self.results = results


class TokenMatchingError(AssertionError):
    pass


# ---- goldens ------------------------------------------------------------------------------------------------
@torch.no_grad()
def generate_expected_logits(hf_model, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor],
                             num_tokens: int, use_cache: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Greedy HF generation on CPU.  -> (logits [num_tokens, B, V] fp32, tokens [B, num_tokens]).
    Rows may be right padded (``attention_mask``); each row is run unpadded so HF sees clean inputs."""
    B = input_ids.shape[0]
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    import inspect
    # ``use_cache=False``: recompute the whole prefix every step (an oracle that does not depend on the HF cache implementation)
    has_cache = use_cache and "past_key_values" in inspect.signature(hf_model.forward).parameters
    all_logits, all_tokens = [], []
    for b in range(B):
        n = int(attention_mask[b].sum())
        seq = input_ids[b:b + 1, :n]
        if not has_cache:                                   # cache-less architectures (GPT-1): recompute the whole prefix every step
            lg, tk = [], []
            for _ in range(num_tokens):
                l = hf_model(seq).logits[0, -1].float()
                lg.append(l)
                tk.append(int(l.argmax()))
                seq = torch.cat([seq, torch.tensor([[tk[-1]]])], 1)
            all_logits.append(torch.stack(lg))
            all_tokens.append(torch.tensor(tk))
            continue
        lg, tk = [], []
        past = None
        cur = seq
        for _ in range(num_tokens):
            out = hf_model(cur, past_key_values=past, use_cache=True)
            if past is None and getattr(out, "past_key_values", None) is None:
                # models whose output type carries no cache (RecurrentGemma): hand in a cache object that is updated in place
                from transformers import DynamicCache
                past = DynamicCache(config=hf_model.config)
                out = hf_model(cur, past_key_values=past, use_cache=True)
            past = getattr(out, "past_key_values", None) or past
            l = out.logits[0, -1].float()
            t = int(l.argmax())
            lg.append(l)
            tk.append(t)
            cur = torch.tensor([[t]])
        all_logits.append(torch.stack(lg))
        all_tokens.append(torch.tensor(tk))
    return torch.stack(all_logits, 1), torch.stack(all_tokens, 0)


# ---- model under test ------------------------------------------------------------------------------------------
@torch.no_grad()
def teacher_forced_logits(model, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor],
                          forced_tokens: torch.Tensor) -> torch.Tensor:
    """Run ``model`` (a NeuronBaseForCausalLM) over prompt + forced continuation.  -> logits [steps, B, V]."""
    B = input_ids.shape[0]
    steps = forced_tokens.shape[1]
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    model.reset()
    out = model(input_ids, attention_mask=attention_mask, output_logits=True)
    logits = [out.logits[:, -1].float().cpu()]
    pos = attention_mask.sum(-1).view(B, 1).to(torch.int32)
    for s in range(steps - 1):
        out = model(forced_tokens[:, s:s + 1], position_ids=pos, output_logits=True)
        logits.append(out.logits[:, -1].float().cpu())
        pos = pos + 1
    return torch.stack(logits, 0)


def _validate_single_token_logits(expected: torch.Tensor, actual: torch.Tensor, tol_map: dict,
                                  divergence_difference_tol: float, remove_shift: bool = True,
                                  actual_token_id: Optional[int] = None) -> Tuple[bool, dict]:
    exp = expected.float()
    act = actual.float()
    V = min(exp.numel(), act.numel())
    exp, act = exp[:V], act[:V]
    res: Dict = {"topk": {}}
    ok = True
    order = torch.argsort(exp, descending=True)
    for key, (atol, rtol) in tol_map.items():
        k = V if key in ("all", None) else min(int(key), V)
        idx = order[:k]
        e, a = exp[idx], act[idx]
        if remove_shift:
            a = a - (a - e).mean()
        err = (a - e).abs()
        lim = atol + rtol * e.abs()
        worst = float((err - lim).max())
        rel = float((err / e.abs().clamp_min(1e-6)).max())
        passed = worst <= 0
        res["topk"][str(key)] = {"passed": passed, "max_rel_err": rel, "max_abs_err": float(err.max())}
        ok &= passed
    exp_tok = int(exp.argmax())
    act_tok = int(act.argmax()) if actual_token_id is None else int(actual_token_id)
    res["expected_token"], res["actual_token"] = exp_tok, act_tok
    if exp_tok != act_tok:
        diff = float((act.max() - act[exp_tok]).abs())
        res["divergence_difference"] = diff
        acceptable = diff <= divergence_difference_tol * max(1.0, float(act.max().abs()))
        res["status"] = ("acceptable_divergence" if acceptable else "diverged") + ("" if ok else "_with_topk_errors")
        ok &= acceptable
    else:
        res["status"] = "matched" if ok else "topk_errors"
    return ok, res


def logit_validation(input_ids: List[List[int]], generate_fn: Callable, expected_logits: torch.Tensor,
                     tol_map: Optional[dict] = None, divergence_difference_tol: Optional[float] = None,
                     suppress_passing: bool = True, colorize: bool = False) -> bool:
    """Generic form (same signature as the reference): ``generate_fn(input_ids)`` returns logits
    ``[steps, B, V]`` (or ``(logits, sequences)``) for a *free-running* generation from ``input_ids``; on a
    divergence the golden prefix is appended to ``input_ids`` and generation restarts from there."""
    tol_map = tol_map or DEFAULT_TOLERANCE_MAP
    dtol = DEFAULT_DIVERGENCE_DIFFERENCE_TOLERANCE if divergence_difference_tol is None else divergence_difference_tol
    B = len(input_ids)
    input_ids = [list(x) for x in input_ids]
    exp_seq = expected_logits.argmax(2).T
    total = exp_seq.shape[1]
    start, passed = 0, True
    results = [[] for _ in range(B)]
    while start < total:
        r = generate_fn(input_ids)
        if isinstance(r, tuple):
            act_logits, act_seq = r
        else:
            act_logits, act_seq = r, r.argmax(2).T
        n = min(act_logits.shape[0], total - start)
        mism = (exp_seq[:, start:start + n] != act_seq[:, :n]).any(0)
        div = int(mism.float().argmax()) + 1 if bool(mism.any()) else n   # include the diverging position
        for b in range(B):
            for t in range(div):
                ok, res = _validate_single_token_logits(expected_logits[start + t, b], act_logits[t, b], tol_map, dtol)
                results[b].append(res)
                passed &= ok
        for b in range(B):
            input_ids[b].extend(exp_seq[b, start:start + div].tolist())
        start += div
    _log_results(results, suppress_passing)
    return passed


def _log_results(results, suppress_passing=True):
    for b, rows in enumerate(results):
        for t, r in enumerate(rows):
            if suppress_passing and r["status"] in ("matched", "acceptable_divergence"):
                continue
            logger.warning("logit validation: batch %d token %d: %s %s", b, t, r["status"],
                           {k: round(v["max_rel_err"], 4) for k, v in r["topk"].items()})


# ---- public checks (reference names) ----------------------------------------------------------------------------
def get_generate_outputs(model, prompts: Sequence[str], tokenizer, is_hf: bool = False, generation_config=None,
                         max_length: Optional[int] = None, **kwargs):
    """Tokenise prompts (right padded), generate, decode.  reference accuracy.py get_generate_outputs."""
    tokenizer.padding_side = "right"
    if tokenizer.pad_token_id is None:
        tokenizer.pad_token = tokenizer.eos_token
    enc = tokenizer(list(prompts), padding=True, return_tensors="pt")
    gen = model if is_hf else HuggingFaceGenerationAdapter(model)
    if is_hf:
        out = gen.generate(enc.input_ids, attention_mask=enc.attention_mask, max_length=max_length, do_sample=False,
                           **kwargs)
    else:
        out = gen.generate(enc.input_ids, attention_mask=enc.attention_mask, generation_config=generation_config,
                           max_length=max_length, **kwargs)
    return out, tokenizer.batch_decode(out, skip_special_tokens=True)


def check_accuracy(model, tokenizer=None, generation_config=None, expected_token_ids: Optional[torch.Tensor] = None,
                   num_tokens_to_check: Optional[int] = None, input_ids: Optional[torch.Tensor] = None,
                   attention_mask: Optional[torch.Tensor] = None, hf_model=None, prompts=None, max_new_tokens=None,
                   **kwargs) -> bool:
    """Token matching against HF-CPU greedy goldens (reference accuracy.py:244-343)."""
    if input_ids is None:
        assert tokenizer is not None and prompts is not None, "need input_ids or (tokenizer, prompts)"
        tokenizer.padding_side = "right"
        if tokenizer.pad_token_id is None:
            tokenizer.pad_token = tokenizer.eos_token
        enc = tokenizer(list(prompts), padding=True, return_tensors="pt")
        input_ids, attention_mask = enc.input_ids, enc.attention_mask
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    nc = model.neuron_config
    n_new = max_new_tokens or (nc.max_length - input_ids.shape[1])
    if num_tokens_to_check:
        n_new = min(n_new, num_tokens_to_check)
    if expected_token_ids is None:
        if hf_model is None:
            hf_model = model.load_hf_model(model.model_path)
        _, expected_token_ids = generate_expected_logits(hf_model, input_ids, attention_mask, n_new)
    adapter = HuggingFaceGenerationAdapter(model)
    out = adapter.generate(input_ids, attention_mask=attention_mask, max_new_tokens=n_new,
                           generation_config=generation_config, **kwargs)
    B = input_ids.shape[0]
    for b in range(B):
        n = int(attention_mask[b].sum())
        got = out[b, n:n + n_new]
        exp = expected_token_ids[b, :n_new]
        m = min(got.numel(), exp.numel())
        if not torch.equal(got[:m].cpu(), exp[:m].cpu()):
            first = int((got[:m].cpu() != exp[:m].cpu()).float().argmax())
            raise TokenMatchingError(f"token mismatch in batch row {b} at generated position {first}: "
                                     f"got {got[:m].tolist()} expected {exp[:m].tolist()}")
    logger.info("token matching passed (%d rows x %d tokens)", B, n_new)
    return True


def check_accuracy_logits(model, tokenizer=None, generation_config=None, expected_logits: Optional[torch.Tensor] = None,
                          divergence_difference_tol: float = DEFAULT_DIVERGENCE_DIFFERENCE_TOLERANCE,
                          tol_map: Optional[dict] = None, num_tokens_to_check: Optional[int] = None,
                          input_ids: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                          hf_model=None, prompts=None, **kwargs):
    """Teacher-forced logit matching (reference accuracy.py:478-704).  Returns the per-token results; raises
    :class:`LogitMatchingValidationError` on failure."""
    if input_ids is None:
        assert tokenizer is not None and prompts is not None
        tokenizer.padding_side = "right"
        if tokenizer.pad_token_id is None:
            tokenizer.pad_token = tokenizer.eos_token
        enc = tokenizer(list(prompts), padding=True, return_tensors="pt")
        input_ids, attention_mask = enc.input_ids, enc.attention_mask
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    nc = model.neuron_config
    steps = num_tokens_to_check or (nc.max_length - input_ids.shape[1])
    if expected_logits is None:
        if hf_model is None:
            hf_model = model.load_hf_model(model.model_path)
        expected_logits, _ = generate_expected_logits(hf_model, input_ids, attention_mask, steps)
    expected_logits = expected_logits[:steps]
    golden = expected_logits.argmax(2).T
    actual = teacher_forced_logits(model, input_ids, attention_mask, golden)
    tol_map = {str(k): tuple(v) for k, v in (tol_map or DEFAULT_TOLERANCE_MAP).items()}
    passed, results = True, [[] for _ in range(input_ids.shape[0])]
    for t in range(actual.shape[0]):
        for b in range(input_ids.shape[0]):
            ok, res = _validate_single_token_logits(expected_logits[t, b], actual[t, b], tol_map, divergence_difference_tol)
            results[b].append(res)
            passed &= ok
    _log_results(results)
    if not passed:
        bad = [(b, t, r["status"]) for b, rows in enumerate(results) for t, r in enumerate(rows)
               if r["status"] not in ("matched", "acceptable_divergence")]
        raise LogitMatchingValidationError(f"logit matching failed at {bad[:8]}", results)
    return results


check_accuracy_logits_v2 = check_accuracy_logits


def check_accuracy_embeddings(actual: torch.Tensor, expected: torch.Tensor, plot_outputs: bool = False,
                              rtol: float = 0.0, atol: float = 0.0) -> Tuple[bool, float]:
    """Element-wise closeness of hidden states / encoder outputs (module tests; reference accuracy.py)."""
    a, e = actual.float().cpu(), expected.float().cpu()
    err = (a - e).abs()
    lim = atol + rtol * e.abs()
    return bool((err <= lim).all()), float(err.max())


def check_draft_logits(draft_logits: torch.Tensor, expected_logits: torch.Tensor, tol_map=None,
                       divergence_difference_tol=DEFAULT_DIVERGENCE_DIFFERENCE_TOLERANCE) -> bool:
    """Speculation: compare the draft model's logits at every speculated position with goldens produced by the
    stand-alone draft model (reference accuracy.py:1222-1277).  Shapes [steps, B, V]."""
    tol_map = tol_map or DEFAULT_TOLERANCE_MAP
    ok = True
    for t in range(min(draft_logits.shape[0], expected_logits.shape[0])):
        for b in range(draft_logits.shape[1]):
            p, _ = _validate_single_token_logits(expected_logits[t, b], draft_logits[t, b], tol_map,
                                                 divergence_difference_tol)
            ok &= p
    return ok


def generate_with_chunked_prefill(neuron_model, input_ids: torch.Tensor, num_tokens: int, chunk_size: Optional[int] = None,
                                  block_table: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Chunked prefill through the paged KV cache, then greedy decode (reference utils/accuracy.py:948-1100): every sequence's prompt
    is encoded ``chunk_size`` tokens at a time — each chunk attends to the blocks written by the previous ones
    (``computed_context_lens``) — and all sequences then decode together.  Block table and slot mapping are generated here
    (sequence ``b`` owns blocks ``[b * n, (b + 1) * n)``) unless a ``block_table`` is given.
    -> logits ``[num_tokens, B, V]`` (first entry: the logits after the last prompt token)."""
    nc = neuron_model.config.neuron_config
    assert nc.is_block_kv_layout, "chunked prefill uses the block KV cache"
    B, T = input_ids.shape
    bs = nc.pa_block_size
    n_blk = -(-(T + num_tokens) // bs)
    if block_table is None:
        assert B * n_blk <= nc.pa_num_blocks, f"need {B * n_blk} KV blocks, have {nc.pa_num_blocks}"
        block_table = (torch.arange(B).view(B, 1) * n_blk + torch.arange(n_blk).view(1, n_blk)).to(torch.int32)
    cpc = getattr(nc, "chunked_prefill_config", None)
    chunk = chunk_size or (getattr(cpc, "kernel_q_tile_size", None) if cpc is not None else None) or nc.max_context_length

    def slots(pos):
        blk = torch.gather(block_table.long(), 1, pos.long() // bs)
        return (blk * bs + pos.long() % bs).to(torch.int32)

    neuron_model.reset()
    out, done = None, 0
    while done < T:
        n = min(chunk, T - done)
        pos = (torch.arange(n) + done).unsqueeze(0).expand(B, n)
        kw = {}
        if done:
            kw = dict(computed_context_lens=torch.full((B,), done), full_context_lens=torch.full((B,), done + n))
        out = neuron_model(input_ids[:, done:done + n], attention_mask=torch.ones(B, n, dtype=torch.int32), position_ids=pos,
                           slot_mapping=slots(pos), block_table=block_table, output_logits=True, **kw)
        done += n
    logits = [out.logits[:, -1].float().cpu()]
    pos = torch.full((B, 1), T, dtype=torch.int32)
    for _ in range(num_tokens - 1):
        tok = logits[-1].argmax(-1).view(B, 1)
        out = neuron_model(tok, position_ids=pos, slot_mapping=slots(pos), block_table=block_table, output_logits=True)
        logits.append(out.logits[:, -1].float().cpu())
        pos = pos + 1
    return torch.stack(logits, 0)
