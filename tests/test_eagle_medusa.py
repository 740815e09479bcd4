"""EAGLE (chain / EAGLE-3 / static token tree) and Medusa speculation: lossless w.r.t. the target's own greedy decoding."""
import pytest
import torch

from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter
from neuronx_distributed_inference_b200.utils.testing import build_random_llama

TINY = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
            vocab_size=96, head_dim=16)
COMMON = dict(batch_size=2, seq_len=64, max_context_length=16, device="cpu", dtype="float32")


def _greedy(ids, mask, n):
    base = build_random_llama(TINY, seed=5, **COMMON)
    return HuggingFaceGenerationAdapter(base).generate(ids, attention_mask=mask, max_new_tokens=n)


def _same(seqs, ref, mask, n_new):
    """Rows of a right-padded batch: the baseline stops ``n_new`` tokens after each row's own prompt, speculation runs every
    row up to the common max length — compare what both have."""
    for b in range(ref.shape[0]):
        n = int(mask[b].sum()) + n_new
        if not torch.equal(seqs[b, :n], ref[b, :n]):
            return False
    return True


def _prompt():
    torch.manual_seed(0)
    ids = torch.randint(1, 96, (2, 7))
    mask = torch.ones_like(ids)
    mask[1, 5:] = 0
    ids[1, 5:] = 0
    return ids, mask


@pytest.mark.parametrize("variant", ["eagle", "eagle3", "eagle_tree"])
def test_eagle_is_lossless(variant):
    ids, mask = _prompt()
    ref = _greedy(ids, mask, 18)
    kw = dict(speculation_length=4, enable_eagle_speculation=True, enable_fused_speculation=True)
    dn = dict(is_eagle_draft=True)
    if variant == "eagle3":
        kw["is_eagle3"] = True
        dn["is_eagle3"] = True
    if variant == "eagle_tree":
        kw["token_tree_config"] = {"0": ["1", "2"], "1": ["3", "4"], "2": ["5"], "3": ["6"]}
    app = build_random_llama(TINY, seed=5, fused_draft=dict(hf=dict(num_hidden_layers=1), neuron=dn), **COMMON, **kw)
    out = HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=18, return_dict_in_generate=True)
    assert _same(out.sequences, ref, mask, 18), (out.sequences, ref)
    assert out.speculation_stats["steps"] > 0


def test_eagle_perfect_features_accept_more_than_one():
    """A draft that *is* the target's last layer + the same head would be perfect; here we only check that acceptance
    statistics are reported and bounded by k per step."""
    ids, mask = _prompt()
    app = build_random_llama(TINY, seed=5, speculation_length=3, enable_eagle_speculation=True, enable_fused_speculation=True,
                             fused_draft=dict(hf=dict(num_hidden_layers=1), neuron=dict(is_eagle_draft=True)), **COMMON)
    out = HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=10, return_dict_in_generate=True)
    st = out.speculation_stats
    assert st["steps"] * 2 <= st["accepted"] * 2 and st["accepted"] <= st["steps"] * 3 * 2


def test_medusa_is_lossless():
    ids, mask = _prompt()
    ref = _greedy(ids, mask, 16)
    app = build_random_llama(TINY, seed=5, is_medusa=True, num_medusa_heads=3, medusa_speculation_length=8,
                             medusa_tree=[[0], [1], [0, 0], [0, 1], [1, 0], [0, 0, 0]], output_logits=True, **COMMON)
    # the medusa heads draw from the same RNG stream after the base weights: base weights stay identical to `ref`'s
    out = HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=16, return_dict_in_generate=True)
    assert _same(out.sequences, ref, mask, 16), (out.sequences, ref)


def test_medusa_buffers_shape():
    from neuronx_distributed_inference_b200.generation.medusa import DEFAULT_MEDUSA_TREE, generate_medusa_buffers
    b = generate_medusa_buffers(DEFAULT_MEDUSA_TREE)
    assert b["medusa_attn_mask"].shape == (64, 64) and b["retrieve_indices"].shape[1] == 5
    assert int(b["medusa_position_ids"].max()) == 4 and b["tree_indices"][0] == 0


def test_tree_verify_accepts_side_branch_and_compacts_kv():
    """Oracle proposals placed on a NON-first branch: the verify step must accept exactly that path, and after KV
    compaction plain decoding must continue as if the tokens had been generated one by one."""
    ids, mask = _prompt()
    ids, mask = ids[:1, :7], mask[:1, :7]
    base = build_random_llama(TINY, seed=5, **dict(COMMON, batch_size=1))
    ref = HuggingFaceGenerationAdapter(base).generate(ids, attention_mask=mask, max_new_tokens=8)[0, 7:].tolist()
    app = build_random_llama(TINY, seed=5, is_medusa=True, num_medusa_heads=3, medusa_speculation_length=8,
                             medusa_tree=[[0], [1], [0, 0], [0, 1], [1, 0], [0, 0, 0]], output_logits=True,
                             **dict(COMMON, batch_size=1))
    med = app.medusa_model
    seq = torch.zeros(1, dtype=torch.int32)
    pos0 = torch.arange(7, dtype=torch.int32).view(1, 7)
    _, root, _ = med.prefill(ids, mask, pos0, seq)
    assert int(root) == ref[0]
    wrong = (ref[1] + 1) % 96
    heads = torch.full((1, 3, 2), wrong, dtype=torch.long)
    heads[0, 0, 1] = ref[1]          # depth-1 truth on rank 1 (node [1])
    heads[0, 1, 0] = ref[2]          # depth-2 truth on rank 0: nodes [0,0] and [1,0] both carry it; only [1,0] is reachable
    acc, n_acc, nxt, _, pos = med(root, heads, torch.tensor([[7]], dtype=torch.int32), seq)
    assert int(n_acc) == 3 and acc[0, :3].tolist() == ref[1:4] and int(nxt) == ref[3] and int(pos) == 10
    # plain decode from the compacted cache
    tok = nxt.view(1, 1)
    got = []
    for i in range(4):
        out = app.model(tok, None, torch.tensor([[10 + i]], dtype=torch.int32), seq, None, is_prefill=False, output_logits=True)
        tok = out.logits[:, -1].argmax(-1).view(1, 1)
        got.append(int(tok))
    assert got == ref[4:8]
