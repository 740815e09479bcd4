"""Speculative decoding: output must equal the target's own greedy decoding, whatever the draft proposes."""
import torch

from neuronx_distributed_inference_b200.generation.speculative import greedy_accept
from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter
from neuronx_distributed_inference_b200.utils.testing import build_random_llama

TINY = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
            vocab_size=128, head_dim=16)


def test_greedy_accept():
    d = torch.tensor([[5, 6, 7], [5, 9, 7], [1, 2, 3]])
    t = torch.tensor([[5, 6, 7, 8], [5, 6, 7, 8], [9, 2, 3, 4]])
    acc, n = greedy_accept(d, t)
    assert n.tolist() == [4, 2, 1]
    assert acc.tolist() == [[5, 6, 7, 8], [5, 6, -1, -1], [9, -1, -1, -1]]


def _plain(app, ids, mask, n):
    return HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=n)


def test_vanilla_speculation_equals_target_greedy():
    target = build_random_llama(TINY, batch_size=2, seq_len=48, max_context_length=16, device="cpu", dtype="float32",
                                seed=1, speculation_length=4)
    draft = build_random_llama(dict(TINY, num_hidden_layers=1), batch_size=2, seq_len=48, max_context_length=16,
                               device="cpu", dtype="float32", seed=2)
    ids = torch.randint(0, 128, (2, 9))
    mask = torch.ones_like(ids)
    mask[1, 6:] = 0
    spec = HuggingFaceGenerationAdapter(target).generate(ids, attention_mask=mask, max_new_tokens=20, assistant_model=draft,
                                                          return_dict_in_generate=True)
    target.neuron_config.speculation_length = 0
    ref = _plain(target, ids, mask, 20)
    for b in range(2):
        n = int(mask[b].sum())
        assert spec.sequences[b, : n + 20].tolist() == ref[b, : n + 20].tolist()
    assert spec.speculation_stats["steps"] <= 2 * 20


def test_perfect_draft_accepts_everything():
    target = build_random_llama(TINY, batch_size=1, seq_len=48, max_context_length=16, device="cpu", dtype="float32",
                                seed=3, speculation_length=5)
    draft = build_random_llama(TINY, batch_size=1, seq_len=48, max_context_length=16, device="cpu", dtype="float32", seed=3)
    ids = torch.randint(0, 128, (1, 7))
    out = HuggingFaceGenerationAdapter(target).generate(ids, max_new_tokens=20, assistant_model=draft,
                                                         return_dict_in_generate=True)
    st = out.speculation_stats
    assert st["accepted"] == 5 * st["steps"]        # identical models: every proposal is accepted
    target.neuron_config.speculation_length = 0
    ref = _plain(target, ids, torch.ones_like(ids), 20)
    assert out.sequences[0, :27].tolist() == ref[0, :27].tolist()


def test_fused_speculation_inside_application():
    """enable_fused_speculation: the application owns the draft (FusedSpecNeuronConfig) — no assistant_model argument."""
    torch.manual_seed(1)
    ids = torch.randint(1, 128, (2, 6))
    base = build_random_llama(TINY, batch_size=2, seq_len=48, max_context_length=16, device="cpu", dtype="float32", seed=3)
    ref = HuggingFaceGenerationAdapter(base).generate(ids, max_new_tokens=12)
    app = build_random_llama(TINY, batch_size=2, seq_len=48, max_context_length=16, device="cpu", dtype="float32", seed=3,
                             speculation_length=3, enable_fused_speculation=True,
                             fused_draft=dict(hf=dict(num_hidden_layers=1)))
    assert app.fused_spec_model is not None and app.draft_model is not None
    out = HuggingFaceGenerationAdapter(app).generate(ids, max_new_tokens=12)
    assert torch.equal(out[:, : ref.shape[1]], ref)


def test_speculative_sampling_follows_the_target_distribution():
    """Rejection-sampling acceptance: whatever the draft proposes, the first emitted token is distributed like the target."""
    from neuronx_distributed_inference_b200.generation.speculative import adjust_target_probs, speculative_sample_accept
    torch.manual_seed(0)
    V, N, k = 6, 60000, 3
    p_t = torch.tensor([0.05, 0.4, 0.1, 0.25, 0.15, 0.05])
    p_d = torch.tensor([0.3, 0.1, 0.3, 0.1, 0.1, 0.1])
    dprobs = p_d.expand(N, k - 1, V).contiguous()
    tprobs = p_t.expand(N, k, V).contiguous()
    dtok = torch.multinomial(p_d, N * (k - 1), replacement=True).view(N, k - 1)
    acc, n_acc = speculative_sample_accept(dtok, dprobs, tprobs, torch.rand(N, k - 1), torch.rand(N, k))
    first = acc[:, 0]
    emp = torch.bincount(first, minlength=V).float() / N
    assert (emp - p_t).abs().max() < 0.01, emp
    assert int(n_acc.min()) >= 1 and int(n_acc.max()) <= k and (acc.gather(1, (n_acc - 1).view(-1, 1)) >= 0).all()
    assert ((acc >= 0).sum(-1) == n_acc).all()
    # identical distributions: every draft token is accepted
    acc2, n2 = speculative_sample_accept(dtok, dprobs, dprobs.new_tensor(p_d).expand(N, k, V).contiguous(), torch.rand(N, k - 1), torch.rand(N, k))
    assert (n2 == k).all() and torch.equal(acc2[:, : k - 1], dtok)
    r = adjust_target_probs(p_t.view(1, V), p_d.view(1, V))
    assert torch.allclose(r.sum(-1), torch.ones(1)) and float(r[0, 0]) == 0.0


def test_fused_speculation_sampling_path_degenerates_to_greedy_and_samples_valid_tokens():
    from neuronx_distributed_inference_b200.config import OnDeviceSamplingConfig
    from neuronx_distributed_inference_b200.modules.sampling import prepare_sampling_params
    torch.manual_seed(2)
    ids = torch.randint(1, 128, (2, 6))
    kw = dict(batch_size=2, seq_len=48, max_context_length=16, device="cpu", dtype="float32", seed=3)
    ref = HuggingFaceGenerationAdapter(build_random_llama(TINY, **kw)).generate(ids, max_new_tokens=10)
    app = build_random_llama(TINY, speculation_length=3, enable_fused_speculation=True, fused_draft=dict(hf=dict(num_hidden_layers=1)),
                             on_device_sampling_config=OnDeviceSamplingConfig(do_sample=True, dynamic=True), **kw)
    ad = HuggingFaceGenerationAdapter(app)
    # top_k = 1: the sampling machinery must reproduce greedy decoding exactly
    out = ad.generate(ids, max_new_tokens=10, sampling_params=prepare_sampling_params(2, 1, 1.0, 1.0))
    assert torch.equal(out[:, : ref.shape[1]], ref)
    # genuine sampling: valid tokens, right length, different draws across seeds are allowed
    out = ad.generate(ids, max_new_tokens=10, sampling_params=prepare_sampling_params(2, 20, 0.9, 1.2))
    assert out.shape[1] >= ids.shape[1] + 10 and int(out.min()) >= 0 and int(out.max()) < 128


def test_fused_speculation_through_application_forward():
    """The reference's calling convention: app.forward returns accepted tokens (padded with -1) and ``fused_outputs``."""
    torch.manual_seed(4)
    ids = torch.randint(1, 128, (2, 6))
    kw = dict(batch_size=2, seq_len=48, max_context_length=16, device="cpu", dtype="float32", seed=3)
    ref = HuggingFaceGenerationAdapter(build_random_llama(TINY, **kw)).generate(ids, max_new_tokens=9)
    app = build_random_llama(TINY, speculation_length=3, enable_fused_speculation=True, fused_draft=dict(hf=dict(num_hidden_layers=1)), **kw)
    out = app(ids, attention_mask=torch.ones_like(ids))
    tok, pos = out.tokens.view(2, 1), torch.full((2, 1), 6, dtype=torch.int32)
    rows = [[int(t)] for t in tok.view(-1)]
    while min(len(r) for r in rows) < 9:
        out = app(tok, position_ids=pos)
        acc, nxt, _, npos, n_acc = out.fused_outputs
        assert out.tokens.shape == (2, 3) and int(n_acc.min()) >= 1
        for b in range(2):
            rows[b] += [t for t in acc[b].tolist() if t >= 0]
        tok, pos = nxt, npos
    for b in range(2):
        assert rows[b][:9] == ref[b, 6:15].tolist()
