"""Serving features at the application level: continuous batching by seq_ids, masked rows, paged KV cache with slot
mapping / block tables, prefix caching, batch larger than the compiled batch, async token feedback flag."""
import torch

from neuronx_distributed_inference_b200.utils.testing import build_random_llama

TINY = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
            vocab_size=128, head_dim=16)


def _greedy(app, ids, steps, seq_ids=None, **kw):
    out = app(ids, attention_mask=torch.ones_like(ids), seq_ids=seq_ids, **kw)
    toks = [out.tokens.clone()]
    pos = torch.full((ids.shape[0], 1), ids.shape[1], dtype=torch.int32)
    for _ in range(steps):
        out = app(toks[-1].view(-1, 1), position_ids=pos, seq_ids=seq_ids)
        toks.append(out.tokens.clone())
        pos = pos + 1
    return torch.stack(toks, 1)


def test_continuous_batching_seq_ids_and_masked_rows():
    app = build_random_llama(TINY, batch_size=3, seq_len=48, max_context_length=16, device="cpu", dtype="float32",
                             is_continuous_batching=True, ctx_batch_size=1, kv_cache_batch_size=3, apply_seq_ids_mask=True)
    ids = torch.randint(0, 128, (3, 8))
    ref = _greedy(app, ids, 5)                     # lines 0,1,2 in order
    app.reset()
    # prefill one sequence at a time into shuffled cache lines, then decode them together in another order
    order = [2, 0, 1]
    first = {}
    for b, line in enumerate(order):
        o = app(ids[b:b + 1], attention_mask=torch.ones(1, 8, dtype=torch.long), seq_ids=torch.tensor([line], dtype=torch.int32))
        first[line] = o.tokens
    seq = torch.tensor([1, 2, 0], dtype=torch.int32)   # rows -> cache lines
    tok = torch.cat([first[int(s)] for s in seq])
    rows = {int(s): [int(first[int(s)])] for s in seq}
    pos = torch.full((3, 1), 8, dtype=torch.int32)
    for _ in range(5):
        o = app(tok.view(3, 1), position_ids=pos, seq_ids=seq)
        tok = o.tokens
        for r, s in enumerate(seq.tolist()):
            rows[s].append(int(tok[r]))
        pos = pos + 1
    for b, line in enumerate(order):
        assert rows[line] == ref[b].tolist()
    # a masked row (seq_id -1) must not disturb the live cache lines
    o1 = app(tok.view(3, 1), position_ids=pos, seq_ids=torch.tensor([1, -1, 0], dtype=torch.int32))
    o2 = app(tok.view(3, 1), position_ids=pos, seq_ids=torch.tensor([1, 2, 0], dtype=torch.int32))
    assert o1.tokens[0] == o2.tokens[0] and o1.tokens[2] == o2.tokens[2]


def test_batch_larger_than_compiled_batch_is_split():
    app = build_random_llama(TINY, batch_size=2, seq_len=32, max_context_length=16, device="cpu", dtype="float32",
                             kv_cache_batch_size=4, max_batch_size=4)
    ids = torch.randint(0, 128, (4, 6))
    seq = torch.arange(4, dtype=torch.int32)
    big = _greedy(app, ids, 3, seq_ids=seq)
    app.reset()
    a = _greedy(app, ids[:2], 3, seq_ids=seq[:2])
    b = _greedy(app, ids[2:], 3, seq_ids=seq[2:])
    assert torch.equal(big, torch.cat([a, b]))


def _slots(block_table, positions, bs):
    blk = torch.gather(block_table.long(), 1, (positions.long() // bs))
    return (blk * bs + positions.long() % bs).int()


def test_paged_kv_matches_contiguous_and_prefix_caching():
    cont = build_random_llama(TINY, batch_size=2, seq_len=64, max_context_length=32, device="cpu", dtype="float32", seed=5)
    paged = build_random_llama(TINY, batch_size=2, seq_len=64, max_context_length=32, device="cpu", dtype="float32", seed=5,
                               is_block_kv_layout=True, pa_block_size=8, pa_num_blocks=24)
    ids = torch.randint(0, 128, (2, 12))
    ref = _greedy(cont, ids, 6)
    bt = torch.tensor([[3, 9, 1, 20, 7, 2, 0, 11], [5, 4, 13, 6, 8, 10, 12, 14]], dtype=torch.int32)
    pos = torch.arange(12).unsqueeze(0).expand(2, 12)
    out = paged(ids, attention_mask=torch.ones_like(ids), position_ids=pos, slot_mapping=_slots(bt, pos, 8), block_table=bt)
    toks = [out.tokens.clone()]
    p = torch.full((2, 1), 12, dtype=torch.int32)
    for _ in range(6):
        out = paged(toks[-1].view(2, 1), position_ids=p, slot_mapping=_slots(bt, p, 8), block_table=bt)
        toks.append(out.tokens.clone())
        p = p + 1
    assert torch.equal(torch.stack(toks, 1), ref)
    # prefix caching: the first 8 tokens are already cached in blocks; encode only the remaining 4
    paged.reset()
    pre = ids[:, :8]
    ppos = torch.arange(8).unsqueeze(0).expand(2, 8)
    paged(pre, attention_mask=torch.ones_like(pre), position_ids=ppos, slot_mapping=_slots(bt, ppos, 8), block_table=bt)
    rest = ids[:, 8:]
    rpos = (torch.arange(4) + 8).unsqueeze(0).expand(2, 4)
    out = paged(rest, attention_mask=torch.ones_like(rest), position_ids=rpos, slot_mapping=_slots(bt, rpos, 8), block_table=bt,
                computed_context_lens=torch.tensor([8, 8]), full_context_lens=torch.tensor([12, 12]))
    assert torch.equal(out.tokens, ref[:, 0])


def test_slot_mapping_generators():
    from neuronx_distributed_inference_b200.modules.kvcache import (generate_fusedspec_slot_mapping,
                                                                   generate_tokengen_slot_mapping, get_active_block_table)
    bt = torch.tensor([[3, 9, 1], [5, 4, 13]], dtype=torch.int32)
    pos = torch.tensor([[9], [17]])
    sm = generate_tokengen_slot_mapping(pos, torch.zeros(2, 1, dtype=torch.int32), bt, 8)
    assert sm.tolist() == [[9 * 8 + 1], [13 * 8 + 1]]
    sm = generate_fusedspec_slot_mapping(torch.tensor([[6], [6]]), torch.zeros(2, 3, dtype=torch.int32), bt, 8, 3)
    assert sm[0].tolist() == [3 * 8 + 6, 3 * 8 + 7, 9 * 8 + 0]
    assert get_active_block_table(bt, torch.tensor([9, 3]), 8).tolist() == [3, 9, 5]


def test_async_session_equals_sync_decode_cpu():
    import torch
    from neuronx_distributed_inference_b200.modules.async_execution import causal_lm_async_execution
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    tiny = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                vocab_size=128, head_dim=16)
    app = build_random_llama(tiny, batch_size=2, seq_len=48, max_context_length=16, device="cpu", dtype="float32", seed=2,
                             async_mode=True)
    ids = torch.randint(1, 128, (2, 6))
    tok = app(ids, attention_mask=torch.ones_like(ids)).tokens.view(2, 1)
    pos = torch.full((2, 1), 6, dtype=torch.int32)
    got = causal_lm_async_execution(app, tok, pos, 5)
    app.reset()
    t = app(ids, attention_mask=torch.ones_like(ids)).tokens.view(2, 1)
    ref = []
    for i in range(5):
        t = app(t, position_ids=pos + i).tokens.view(2, 1)
        ref.append(t.view(-1))
    assert torch.equal(got, torch.stack(ref, 1))


def test_windowed_context_encoding_equals_single_shot():
    import torch
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    tiny = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                vocab_size=128, head_dim=16)
    kw = dict(batch_size=2, seq_len=64, max_context_length=32, device="cpu", dtype="float32", seed=2, output_logits=True)
    ids = torch.randint(1, 128, (2, 20))
    mask = torch.ones_like(ids)
    mask[1, 7:] = 0            # row 1 ends inside the first window
    ref_app = build_random_llama(tiny, **kw)
    ref = ref_app(ids, attention_mask=mask)
    app = build_random_llama(tiny, windowed_context_encoding_size=8, **kw)
    out = app(ids, attention_mask=mask)
    assert torch.allclose(out.logits, ref.logits, atol=1e-4) and torch.equal(out.tokens, ref.tokens)
    pos = mask.sum(-1).view(2, 1).int()
    a = app(out.tokens.view(2, 1), position_ids=pos)
    b = ref_app(ref.tokens.view(2, 1), position_ids=pos)
    assert torch.allclose(a.logits, b.logits, atol=1e-4)


def test_generate_with_chunked_prefill_matches_one_shot_prefill():
    """reference utils/accuracy.py:948-1100: prompts encoded 5 tokens at a time through the paged cache (each chunk reads the blocks
    the earlier chunks wrote) give the same logits as one-shot prefill + decode."""
    from neuronx_distributed_inference_b200.utils.accuracy import generate_with_chunked_prefill
    kw = dict(batch_size=2, seq_len=64, max_context_length=32, device="cpu", dtype="float32", seed=5, output_logits=True)
    cont = build_random_llama(TINY, **kw)
    paged = build_random_llama(TINY, is_block_kv_layout=True, pa_block_size=8, pa_num_blocks=24, **kw)
    ids = torch.randint(0, 128, (2, 13))
    got = generate_with_chunked_prefill(paged, ids, num_tokens=5, chunk_size=5)
    out = cont(ids, attention_mask=torch.ones_like(ids))
    exp = [out.logits[:, -1].float()]
    pos = torch.full((2, 1), 13, dtype=torch.int32)
    for _ in range(4):
        out = cont(exp[-1].argmax(-1).view(2, 1), position_ids=pos)
        exp.append(out.logits[:, -1].float())
        pos = pos + 1
    exp = torch.stack(exp, 0)
    assert got.shape == exp.shape and ((got - exp).norm() / exp.norm()) < 1e-5


def test_sampling_param_inference_and_dp_validation():
    import pytest
    from neuronx_distributed_inference_b200.modules.attention.utils import validate_tp_prefill_to_dp_decode
    from neuronx_distributed_inference_b200.modules.generation.sampling import infer_sampling_params, prepare_sampling_params, rand_like
    p = infer_sampling_params(prepare_sampling_params(3, [5, 40, 1], [0.9, 0.5, 1.0], [0.7, 0.0, 1.0]))
    assert p.tolist()[1] == [1.0, 1.0, 1.0] and p[0].tolist() == pytest.approx([5.0, 0.9, 0.7])
    r = rand_like(torch.zeros(4, 7))
    assert r.shape == (4, 7) and (r >= 0).all() and (r < 1).all()
    validate_tp_prefill_to_dp_decode(num_kv_heads=8, world_size=32, dp_degree=4)      # 4 ranks replicate each KV head
    with pytest.raises(ValueError):
        validate_tp_prefill_to_dp_decode(num_kv_heads=8, world_size=8, dp_degree=2)


def test_continuous_batching_scheduler_example_matches_isolated_generation():
    """examples/continuous_batching_demo.py: staggered arrivals, more requests than cache lines, line reuse, masked idle rows — every
    request must produce exactly what it produces when it runs alone."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("cb_demo", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                            "examples", "continuous_batching_demo.py"))
    demo = importlib.util.module_from_spec(spec)
    sys_modules_guard = spec.name
    import sys
    sys.modules[sys_modules_guard] = demo
    spec.loader.exec_module(demo)
    kw = dict(seq_len=64, max_context_length=16, device="cpu", dtype="float32", seed=9)
    app = build_random_llama(TINY, batch_size=3, is_continuous_batching=True, ctx_batch_size=1, kv_cache_batch_size=3, apply_seq_ids_mask=True, **kw)
    g = torch.Generator().manual_seed(1)
    reqs = [demo.Request(i, torch.randint(1, 128, (int(torch.randint(3, 12, (1,), generator=g)),), generator=g).tolist(),
                         int(torch.randint(2, 9, (1,), generator=g)), arrival_step=i) for i in range(7)]
    done = demo.ContinuousBatcher(app).run(reqs)
    assert [r.rid for r in done] == list(range(7)) and len({r.line for r in done}) <= 3
    solo = build_random_llama(TINY, batch_size=1, **kw)
    for r in done:
        solo.reset()
        ids = torch.tensor([r.prompt])
        exp = _greedy(solo, ids, r.max_new_tokens - 1)[0].tolist() if r.max_new_tokens > 1 else [int(solo(ids, attention_mask=torch.ones_like(ids)).tokens)]
        assert r.output == exp[: r.max_new_tokens], r.rid


def test_hf_adapter_is_a_generation_mixin_and_runs_hf_machinery():
    """``HuggingFaceGenerationAdapter`` is a ``PreTrainedModel + GenerationMixin`` (reference hf_adapter.py:104): plain greedy through
    ``GenerationMixin.generate`` -> our ``_sample`` equals the lean host loop; HF logits processors (min_new_tokens, repetition penalty,
    a user ``LogitsProcessor``), ``StoppingCriteriaList`` objects and streamers are honoured."""
    from transformers import GenerationConfig, LogitsProcessor, LogitsProcessorList, PreTrainedModel, StoppingCriteria, StoppingCriteriaList
    from transformers.generation.utils import GenerationMixin
    from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    app = build_random_llama(dict(vocab_size=160), batch_size=2, seq_len=64, max_context_length=16, device="cpu", dtype="float32",
                             output_logits=True, seed=3)
    ad = HuggingFaceGenerationAdapter(app)
    assert isinstance(ad, PreTrainedModel) and isinstance(ad, GenerationMixin) and ad.can_generate()
    ids = torch.randint(1, 160, (2, 7))
    mask = torch.ones_like(ids)
    lean = ad.generate(ids, attention_mask=mask, max_new_tokens=10)
    hf = ad.generate(ids, attention_mask=mask, max_new_tokens=10, use_hf_generate=True, do_sample=False)
    assert torch.equal(lean, hf)

    class Ban(LogitsProcessor):                      # forbid the token greedy decoding would pick first
        def __init__(self, tok):
            self.tok = tok

        def __call__(self, input_ids, scores):
            scores[:, self.tok] = float("-inf")
            return scores
    banned = int(lean[0, 7])
    out = ad.generate(ids, attention_mask=mask, max_new_tokens=6, logits_processor=LogitsProcessorList([Ban(banned)]), do_sample=False)
    assert banned not in out[:, 7:].tolist()[0] and out.shape == (2, 13)

    class StopAt(StoppingCriteria):
        def __call__(self, input_ids, scores, **kw):
            return torch.full((input_ids.shape[0],), input_ids.shape[1] >= 10, dtype=torch.bool)
    out = ad.generate(ids, attention_mask=mask, max_new_tokens=20, stopping_criteria=StoppingCriteriaList([StopAt()]), do_sample=False)
    assert out.shape[1] == 10 and torch.equal(out, lean[:, :10])

    eos = int(lean[0, 8])                            # min_new_tokens keeps EOS away until 5 tokens are out
    gc = GenerationConfig(max_new_tokens=8, min_new_tokens=5, eos_token_id=eos, pad_token_id=0, do_sample=False)
    out = ad.generate(ids, attention_mask=mask, generation_config=gc)
    assert eos not in out[0, 7:12].tolist()

    class Collect:
        def __init__(self):
            self.items, self.ended = [], False

        def put(self, v):
            self.items.append(v.clone())

        def end(self):
            self.ended = True
    st = Collect()
    out = ad.generate(ids[:1], attention_mask=mask[:1], max_new_tokens=4, streamer=st, do_sample=False, return_dict_in_generate=True,
                      output_scores=True)
    # (HF's generate() hands the prompt to the streamer first, then one item per new token)
    assert st.ended and len(st.items) == 5 and torch.equal(st.items[0], ids[:1])
    assert torch.equal(torch.cat(st.items[1:]), out.sequences[0, 7:]) and len(out.scores) == 4
