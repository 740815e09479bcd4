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
