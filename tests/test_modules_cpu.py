"""Module-level unit tests mirroring the reference's test/unit inventory (kvcache/*, eagle/*, flashdecode/*, generation/*,
attention utils, moe init, utils/*): pure-PyTorch logic on CPU."""
import pytest
import torch

from neuronx_distributed_inference_b200.ops import reference as ref


# ---- kvcache (reference test/unit/modules/kvcache/*) ------------------------------------------------------------------
def test_kv_cache_manager_update_garbage_line_and_reset():
    from neuronx_distributed_inference_b200.modules.kvcache import KVCacheManager
    m = KVCacheManager(num_layers=2, num_kv_heads=2, head_dim=8, max_len=16, num_lines=3, dtype=torch.float32)
    assert len(m.past_key_values) == 4 and m.past_key_values[0].shape == (4, 2, 16, 8)      # 3 lines + garbage
    k, v = torch.randn(2, 3, 2, 8), torch.randn(2, 3, 2, 8)
    seq = torch.tensor([2, -1])                    # second row masked -> garbage line
    pos = torch.tensor([[4, 5, -1], [0, 1, 2]])    # -1 = padding, skipped
    m.update(1, k, v, seq, pos)
    kc, _ = m.get_kv_by_layer_id(1)
    assert torch.equal(kc[2, :, 4], k[0, 0]) and torch.equal(kc[2, :, 5], k[0, 1]) and kc[2, :, 6].abs().sum() == 0
    assert kc[:3].abs().sum() == k[0, :2].abs().sum()          # nothing of row 1 reached a real line
    m.reset()
    assert m.cache.abs().sum() == 0


def test_data_parallel_kv_manager_remaps_and_masks_foreign_sequences():
    from neuronx_distributed_inference_b200.modules.kvcache import DataParallelKVCacheManager
    m = DataParallelKVCacheManager(num_layers=1, num_kv_heads=1, head_dim=4, max_len=8, num_lines=2, dtype=torch.float32, dp_rank=1, dp_size=2)
    lines = m.lines_for(torch.tensor([0, 1, 2, 3]))
    assert lines.tolist() == [2, 2, 0, 1]                        # seq 0/1 belong to dp rank 0 -> garbage (line 2)


def test_block_kv_manager_and_slot_helpers():
    from neuronx_distributed_inference_b200.modules.kvcache import (BlockKVCacheManager, generate_fusedspec_slot_mapping,
                                                                    generate_tokengen_slot_mapping, get_active_block_table)
    m = BlockKVCacheManager(num_layers=1, num_kv_heads=2, head_dim=4, num_blocks=4, block_size=4, dtype=torch.float32)
    k, v = torch.randn(1, 3, 2, 4), torch.randn(1, 3, 2, 4)
    m.update(0, k, v, torch.tensor([[5, 6, -1]]))
    kc, _ = m.get_kv_by_layer_id(0)
    assert torch.equal(kc[1, 1], k[0, 0]) and torch.equal(kc[1, 2], k[0, 1])
    bt = torch.tensor([[2, 0, 3]])
    assert generate_tokengen_slot_mapping(torch.tensor([[5]]), torch.zeros(1, 1, dtype=torch.long), bt, 4).tolist() == [[0 * 4 + 1]]
    assert generate_fusedspec_slot_mapping(torch.tensor([[3]]), torch.zeros(1, 1, dtype=torch.long), bt, 4, 3).tolist() == [[11, 0, 1]]
    assert get_active_block_table(bt, torch.tensor([6]), 4).tolist() == [2, 0]


def test_kv_cache_fp8_quantization_roundtrip():
    from neuronx_distributed_inference_b200.config import KVQuantizationConfig
    from neuronx_distributed_inference_b200.modules.kvcache import KVCacheManager
    q = KVQuantizationConfig(dtype="float8_e4m3fn", scale_mode="static", k_scale=0.05, v_scale=0.05)
    m = KVCacheManager(num_layers=1, num_kv_heads=1, head_dim=8, max_len=4, num_lines=1, dtype=torch.float32, quant_config=q)
    k = torch.randn(1, 1, 1, 8)
    m.update(0, k, k, torch.tensor([0]), torch.tensor([[0]]))
    kc, _ = m.get_kv_by_layer_id(0)
    assert kc.dtype == torch.float8_e4m3fn
    assert (kc[0, :, 0].float() * 0.05 - k[0, 0]).abs().max() < 0.1


# ---- eagle (reference test/unit/modules/eagle/*) --------------------------------------------------------------------------
def test_hidden_state_rolling_buffer_wraps_and_masks():
    from neuronx_distributed_inference_b200.modules.eagle.hidden_state import HiddenStateRollingBuffer, TokenRollingBuffer
    b = HiddenStateRollingBuffer(2, 4, 3, torch.float32)
    h = torch.arange(2 * 2 * 3, dtype=torch.float32).view(2, 2, 3)
    b.set_state(torch.tensor([0, 1]), torch.tensor([[3, 4], [9, 10]]), h)
    assert torch.equal(b.get_state(torch.tensor([0]), torch.tensor([[4]]))[0, 0], h[0, 1])      # 4 % 4 == 0 slot
    assert torch.equal(b.hidden_states[0, 0], h[0, 1]) and torch.equal(b.hidden_states[1, 1], h[1, 0])
    b.set_state(torch.tensor([-1, 5]), torch.tensor([[0, 1], [0, 1]]), torch.ones(2, 2, 3))       # invalid ids -> garbage row
    assert b.hidden_states[:2].sum() == h.sum()
    t = TokenRollingBuffer(1, 4)
    t.set_tokens(torch.tensor([0]), torch.tensor([[6]]), torch.tensor([[42]]))
    assert int(t.get_tokens(torch.tensor([0]), torch.tensor([[6]]))) == 42


def test_static_token_tree_structures():
    from neuronx_distributed_inference_b200.modules.eagle.token_tree import TokenTree
    t = TokenTree({"0": ["1", "2"], "1": ["3", "4"], "2": ["5"], "5": ["6"]})
    assert t.num_nodes == 7 and t.max_depth == 3 and t.level_width == [1, 2, 3, 1]
    assert t.parent == [-1, 0, 0, 1, 1, 2, 5] and t.child_rank == [0, 0, 1, 0, 1, 0, 0]
    assert t.attn_mask[6].tolist() == [True, False, True, False, False, True, True]              # 6 sees 0,2,5,6
    assert t.position_offsets.tolist() == [0, 1, 1, 2, 2, 2, 3] and t.paths.shape == (3, 4)
    assert t.level_mask(2).shape == (3, 7)
    with pytest.raises(ValueError):
        TokenTree({"0": ["1"], "2": ["1"]})            # node with two parents / two roots
    # Medusa path-list format
    m = TokenTree([[0], [0, 0], [1], [0, 1], [2]])
    assert m.num_nodes == 6 and m.child_rank[1:4] == [0, 1, 2]


def test_dynamic_token_tree_selection_is_ancestor_closed_and_accepts_best_path():
    from neuronx_distributed_inference_b200.modules.eagle.dynamic_token_tree import DynamicTokenTree
    torch.manual_seed(0)
    d = DynamicTokenTree(steps=3, branching_factor=3, step_width=2, num_verify=6)
    st = d.init_state(torch.tensor([5, 9]))
    f = torch.zeros(2, 1, dtype=torch.long)
    for _ in range(3):
        f = d.expand(st, f, torch.randn(2, f.shape[1], 50).log_softmax(-1))
    sel, tok, dep, mask = d.select(st)
    assert sel.shape == (2, 6) and (sel[:, 0] == 0).all() and (dep[:, 0] == 0).all()
    for b in range(2):                                   # every selected node's parent is selected too
        ids = set(sel[b].tolist())
        for n in sel[b, 1:].tolist():
            assert int(st["parent"][b, n]) in ids
    assert mask[:, :, 0].all() and torch.equal(mask.diagonal(dim1=1, dim2=2), torch.ones(2, 6, dtype=torch.bool))
    # target agrees with the first child of the root only
    tgt = torch.zeros(2, 6, dtype=torch.long)
    child = [int((dep[b] == 1).nonzero()[0]) for b in range(2)]
    for b in range(2):
        tgt[b, 0] = tok[b, child[b]]
        tgt[b, child[b]] = -7
    path, n_acc, acc = DynamicTokenTree.accept(tok, dep, mask, tgt)
    assert n_acc.tolist() == [2, 2] and path[:, 1].tolist() == child and (acc[:, 1] == -7).all()


# ---- flash decoding utils (reference test/unit/modules/flashdecode/*) -------------------------------------------------------
def test_flashdecode_slots_horizon_and_combine_equal_full_attention():
    from neuronx_distributed_inference_b200.modules import flashdecode as fd
    from neuronx_distributed_inference_b200.parallel.state import Group
    assert fd.calculate_num_cores_per_group(32, 8, 32) == 4 and fd.calculate_num_cores_per_group(32, 8, 8) == 1
    pos = torch.tensor([[0, 5, 6, -1]])
    assert fd.local_slots(pos, 1, 2).tolist() == [[-1, 2, -1, -1]] and fd.local_slots(pos, 0, 2).tolist() == [[0, -1, 3, -1]]
    assert fd.local_horizon(torch.tensor([[0, 5]]), 1, 2).tolist() == [[-1, 2]]
    # two shards of a sequence merged == attention over the whole sequence
    torch.manual_seed(0)
    S, r, D = 10, 2, 8
    q, k, v = torch.randn(1, 1, 2, D), torch.randn(1, 1, S, D), torch.randn(1, 1, S, D)
    P = torch.tensor([[7]])
    parts = []
    for j in range(r):
        ks, vs = k[:, :, j::r], v[:, :, j::r]
        parts.append(fd.partial_attention(q, ks, vs, fd.local_horizon(P, j, r), D ** -0.5))
    O, M, L = (torch.stack(x) for x in zip(*parts))
    gm = M.amax(0)
    w = torch.exp(M - gm)
    merged = (O * w.unsqueeze(-1)).sum(0) / (L * w).sum(0).unsqueeze(-1)
    full = ref.attention_decode(q, k, v, torch.tensor([0]), P, D ** -0.5)
    assert torch.allclose(merged, full.float(), atol=1e-5)
    assert torch.allclose(fd.combine(*parts[0], Group([0])), parts[0][0] / parts[0][2].unsqueeze(-1))


# ---- generation (reference test/unit/modules/generation/*) ---------------------------------------------------------------------
def test_mask_padded_logits_and_sampler_determinism():
    from neuronx_distributed_inference_b200.config import NeuronConfig, OnDeviceSamplingConfig
    from neuronx_distributed_inference_b200.modules.sampling import Sampler, mask_padded_logits, prepare_sampling_params
    x = torch.zeros(2, 8)
    assert torch.equal(mask_padded_logits(x, 0, 2, 3), x)                    # only the last rank holds the pad columns
    m = mask_padded_logits(x, 1, 2, 3)
    assert (m[:, 5:] < -1e30).all() and (m[:, :5] == 0).all()
    nc = NeuronConfig(on_cpu=True, on_device_sampling_config=OnDeviceSamplingConfig(do_sample=True, dynamic=True, deterministic=True))
    s = Sampler(nc)
    logits = torch.randn(3, 100)
    p = prepare_sampling_params(3, [1, 10, 0], [1.0, 0.8, 1.0], [1.0, 0.7, 1.5])
    a, b = s(logits, p), s(logits, p)
    assert torch.equal(a, b) and int(a[0]) == int(logits[0].argmax())        # deterministic; top_k=1 row is greedy
    nc2 = NeuronConfig(on_cpu=True, on_device_sampling_config=OnDeviceSamplingConfig(do_sample=True, dynamic=True, seed=3))
    draws = {int(Sampler(nc2)(torch.zeros(1, 50), prepare_sampling_params(1, 0, 1.0, 1.0))) for _ in range(1)}
    assert len(draws) == 1


def test_topk_sampling_respects_k_and_p():
    B, V = 64, 40
    logits = torch.randn(B, V)
    top_k = torch.full((B,), 3, dtype=torch.int32)
    tok = ref.sample(logits, top_k, torch.ones(B), torch.ones(B), torch.rand(B))
    top3 = logits.topk(3, -1).indices
    assert (tok.view(B, 1) == top3).any(-1).all()
    # top_p -> 0+ keeps only the arg-max
    tok = ref.sample(logits, torch.zeros(B, dtype=torch.int32), torch.full((B,), 1e-6), torch.ones(B), torch.rand(B))
    assert torch.equal(tok, logits.argmax(-1))


# ---- attention utils / masks (reference test/unit/modules/attention/*) --------------------------------------------------------
def test_mask_builders_and_rope_variants():
    m = ref.build_mask(torch.tensor([[3, 4]]), 6)[0, 0]
    assert m.int().tolist() == [[1, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 0]]
    assert ref.build_mask(torch.tensor([[4]]), 6, window=2)[0, 0, 0].int().tolist() == [0, 0, 0, 1, 1, 0]
    assert ref.build_mask(torch.tensor([[4]]), 6, chunk=3)[0, 0, 0].int().tolist() == [0, 0, 0, 1, 1, 0]
    x = torch.randn(1, 2, 1, 8)
    cos, sin = torch.ones(1, 2, 4), torch.zeros(1, 2, 4)
    assert torch.allclose(ref.apply_rope(x, cos, sin, False), x) and torch.allclose(ref.apply_rope(x, cos, sin, True), x)
    from neuronx_distributed_inference_b200.modules.rope import MRotaryEmbedding, RotaryEmbedding
    r = RotaryEmbedding(8, 64, 10000.0)
    c, s = r(torch.tensor([[0, 5]]))
    assert c.shape == (1, 2, 4) and torch.allclose(c[0, 0], torch.ones(4)) and torch.allclose(s[0, 0], torch.zeros(4))
    mr = MRotaryEmbedding(8, 64, 10000.0, [1, 1, 2])
    p3 = torch.tensor([[[1]], [[2]], [[3]]])
    c3, _ = mr(p3)
    assert torch.allclose(c3[0, 0, 0], r(torch.tensor([[1]]))[0][0, 0, 0]) and torch.allclose(c3[0, 0, 1], r(torch.tensor([[2]]))[0][0, 0, 1])
    assert torch.allclose(c3[0, 0, 2:], r(torch.tensor([[3]]))[0][0, 0, 2:])


# ---- MoE glue (reference test/unit/modules/test_init_moe_module.py) -----------------------------------------------------------
def test_router_and_expert_mlps_match_dense_reference():
    logits = torch.tensor([[2.0, 1.0, 0.0, -1.0]])
    w, i = ref.moe_route(logits, 2, "softmax", True)
    assert i.tolist() == [[0, 1]] and torch.allclose(w.sum(-1), torch.ones(1))
    w2, _ = ref.moe_route(logits, 2, "softmax", False)
    assert torch.allclose(w2, logits.softmax(-1)[:, :2])
    torch.manual_seed(0)
    E, H, I = 4, 16, 8
    x = torch.randn(3, H)
    wgu, wd = torch.randn(E, 2 * I, H) * 0.2, torch.randn(E, H, I) * 0.2
    tw, ti = torch.rand(3, 2), torch.tensor([[0, 3], [1, 1], [2, 0]])
    got = ref.moe_experts(x, wgu, wd, tw, ti)
    exp = torch.zeros(3, H)
    for t in range(3):
        for j in range(2):
            e = int(ti[t, j])
            gu = wgu[e] @ x[t]
            exp[t] += tw[t, j] * (wd[e] @ (torch.nn.functional.silu(gu[:I]) * gu[I:]))
    # a token that picks the same expert twice sums both affinities
    assert torch.allclose(got, exp, atol=1e-5)


# ---- utils (reference test/unit/utils/*) ---------------------------------------------------------------------------------------
def test_argparse_version_distributed_random_utils():
    import argparse
    from neuronx_distributed_inference_b200.utils import argparse_utils as au, distributed as du, version_utils as vu
    from neuronx_distributed_inference_b200.utils.random import set_random_seed
    p = argparse.ArgumentParser()
    p.add_argument("--kv", nargs="+", action=au.KeyValueDict)
    p.add_argument("--mix", nargs="+", action=au.StringOrIntegers)
    a = p.parse_args(["--kv", "a=1", "b=[1,2]", "c=x", "--mix", "3", "-4", "z"])
    assert a.kv == {"a": 1, "b": [1, 2], "c": "x"} and a.mix == [3, -4, "z"]
    assert au.comma_ints("1, 2,3") == [1, 2, 3] and au.json_or_path('{"a": 1}') == {"a": 1}
    assert du.get_rank() == 0 and du.get_world_size() == 1 and du.is_rank_zero()
    assert vu.get_torch_version() >= (2, 0)
    set_random_seed(7)
    x = torch.rand(3)
    set_random_seed(7)
    assert torch.equal(x, torch.rand(3))


def test_input_truncation_and_bucket_overflow_errors():
    """Prompts longer than the largest context bucket are rejected loudly (reference models/test_input_truncation.py)."""
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    tiny = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=1, vocab_size=64,
                head_dim=16)
    app = build_random_llama(tiny, batch_size=1, seq_len=32, max_context_length=8, device="cpu", dtype="float32")
    with pytest.raises((ValueError, AssertionError, RuntimeError)):
        app(torch.randint(1, 64, (1, 12)), attention_mask=torch.ones(1, 12, dtype=torch.long))
    out = app(torch.randint(1, 64, (1, 5)), attention_mask=torch.ones(1, 5, dtype=torch.long))
    assert out.tokens.shape == (1,)


def test_hybrid_kv_cache_manager_rolling_slots():
    """Sliding layers live in a window-sized allocation written modulo the window (reference gpt_oss_kv_cache_manager.py:30-396)."""
    from neuronx_distributed_inference_b200.modules.kvcache.gpt_oss_kv_cache_manager import GptOssKVCacheManager, HybridKVCacheManager
    assert GptOssKVCacheManager is HybridKVCacheManager
    m = HybridKVCacheManager([4, None, 4], num_kv_heads=2, head_dim=8, max_len=32, num_lines=2, dtype=torch.float32)
    assert m.rolling_window(0) == 4 and m.rolling_window(1) is None
    assert m.get_kv_by_layer_id(0)[0].shape == (3, 2, 4, 8) and m.get_kv_by_layer_id(1)[0].shape == (3, 2, 32, 8)
    assert len(m.past_key_values) == 6 and m.bytes() == (2 * 2 * 3 * 2 * 4 * 8 + 2 * 3 * 2 * 32 * 8) * 4
    # prefill of 7 tokens (row 1 has 5 valid): only the last 4 positions of each row are written, at pos % 4
    wp = torch.tensor([[0, 1, 2, 3, 4, 5, 6], [0, 1, 2, 3, 4, -1, -1]])
    slots, hor = m.rolling_positions(4, wp, wp.clamp(min=0), True)
    assert slots.tolist() == [[-1, -1, -1, 3, 0, 1, 2], [-1, 1, 2, 3, 0, -1, -1]]
    k = torch.arange(2 * 7, dtype=torch.float32).view(2, 7, 1, 1).expand(2, 7, 2, 8).contiguous()
    m.update(0, k, k, torch.tensor([0, 1]), slots)
    kc = m.get_kv_by_layer_id(0)[0]
    assert kc[0, 0, :, 0].tolist() == [4.0, 5.0, 6.0, 3.0] and kc[1, 0, :, 0].tolist() == [11.0, 8.0, 9.0, 10.0]
    # decode: slot = pos % W, horizon saturates at W - 1
    slots, hor = m.rolling_positions(4, torch.tensor([[7], [2]]), torch.tensor([[7], [2]]), False)
    assert slots.tolist() == [[3], [2]] and hor.tolist() == [[3], [2]]
    with pytest.raises(NotImplementedError):
        m.rolling_positions(4, torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 2, dtype=torch.long), False)
    with pytest.raises(NotImplementedError):
        HybridKVCacheManager([4, 8], 2, 8, 32, 2)


def test_multimodal_kv_cache_manager_vision_lines():
    """Vision K/V are written once per cache line and read back by line at decode (reference multimodal_kv_cache_manager.py:11-134)."""
    from neuronx_distributed_inference_b200.modules.kvcache.multimodal_kv_cache_manager import MultimodalKVCacheManager
    m = MultimodalKVCacheManager(4, 2, 8, 16, 3, torch.float32, cross_attention_layers=[1, 3])
    assert not m.has_vision(1)
    k, v = torch.randn(2, 2, 5, 8), torch.randn(2, 2, 5, 8)
    rm = torch.tensor([[1, 1, 0, 0, 0], [1, 1, 1, 1, 0]], dtype=torch.bool)
    m.update_vision(1, torch.tensor([2, 0]), k, v, rm)
    assert m.has_vision(1) and not m.has_vision(3)
    k2, v2, rm2 = m.get_vision(1, torch.tensor([0, 2]))
    assert torch.equal(k2, k.flip(0)) and torch.equal(v2, v.flip(0)) and torch.equal(rm2, rm.flip(0))
    assert m.bytes() > KVBYTES(m)
    m.reset()
    assert not m.has_vision(1)


def KVBYTES(m):
    return sum(b.numel() * b.element_size() for b in m.buffers())


def test_recurrent_state_cache_lines_and_garbage():
    from neuronx_distributed_inference_b200.modules.kvcache.recurrent_state_cache import RecurrentStateCache
    c = RecurrentStateCache({"conv0": (2, 4), "lru0": (4,)}, num_lines=3, dtype=torch.float32)
    c.write("conv0", torch.tensor([2, 0]), torch.arange(16.0).view(2, 2, 4))
    c.write("lru0", torch.tensor([1, -1]), torch.ones(2, 4))            # -1 -> garbage line, never read back by a live row
    assert torch.equal(c.read("conv0", torch.tensor([0, 2])), torch.arange(16.0).view(2, 2, 4).flip(0))
    assert c.read("lru0", torch.tensor([1]))[0].sum() == 4 and c.read("lru0", torch.tensor([0]))[0].sum() == 0
    assert c.bytes() == (4 * 2 * 4 + 4 * 4) * 4
    c.reset()
    assert c.read("conv0", torch.tensor([2])).abs().sum() == 0


def test_moe_v2_config_objects_drive_expert_mlps():
    """RoutedExpertsMLPOpsConfig (GLU flavour, activation scale / bias, clamps) and MoEFusedTKGConfig (kernel switches) on
    ExpertMLPsV2 (reference moe_v2.py:23-128)."""
    from neuronx_distributed_inference_b200.modules.moe_v2 import (BlockwiseMatmulConfig, ExpertMLPsV2, MoEFusedTKGConfig,
                                                                  RoutedExpertsMLPOpsConfig)
    torch.manual_seed(0)
    E, H, I, k = 4, 16, 8, 2
    rc = RoutedExpertsMLPOpsConfig(num_experts=E, hidden_size=H, intermediate_size=I, top_k=k, glu_type="swiglu", hidden_act_scaling_factor=1.702,
                                   hidden_act_bias=1.0, gate_clamp_upper_limit=7.0, up_clamp_upper_limit=7.0, up_clamp_lower_limit=-7.0)
    assert RoutedExpertsMLPOpsConfig(E, H, I, k).activation() is None            # plain SwiGLU stays on the fused epilogue
    m = ExpertMLPsV2(rc, BlockwiseMatmulConfig(block_size=256), tkg_config=MoEFusedTKGConfig(expert_mlp_kernel_enabled=False))
    m.gate_up_proj.copy_(torch.randn(E, 2 * I, H) * 3)
    m.down_proj.copy_(torch.randn(E, H, I))
    x = torch.randn(5, H)
    w, idx = torch.rand(5, k), torch.randint(0, E, (5, k))
    y = m(x, w, idx)
    ref = torch.zeros(5, H)
    for n in range(5):
        for j in range(k):
            e = int(idx[n, j])
            g, u = (m.gate_up_proj[e] @ x[n]).chunk(2)
            g, u = g.clamp(max=7.0), u.clamp(-7.0, 7.0)
            ref[n] += w[n, j] * (m.down_proj[e] @ ((u + 1.0) * g * torch.sigmoid(1.702 * g)))
    assert torch.allclose(y, ref, atol=1e-4, rtol=1e-4)
    assert m.blockwise_matmul_config.block_size == 256 and not m.tkg_config.decode_kernel_allowed()
    assert BlockwiseMatmulConfig.from_kwargs(block_size=128, not_a_field=1).block_size == 128


def test_kvcache_utils_reference_names():
    from neuronx_distributed_inference_b200.modules.kvcache import utils as U
    c = torch.zeros(3, 2, 8, 4)
    U.fill_prefix(c, torch.ones(2, 2, 3, 4))
    assert c[:2, :, :3].sum() == 2 * 2 * 3 * 4 and c[2].sum() == 0 and c[:, :, 3:].sum() == 0
    t = torch.zeros(4, 6)
    U.dynamic_update_slice(t, torch.ones(2, 3), [3, 5])                       # clamped to fit: rows 2-3, cols 3-5
    assert t[2:, 3:].sum() == 6 and t.sum() == 6
    c = torch.zeros(3, 2, 8, 4)
    U.update_cache_const_indices(c, torch.full((2, 2, 5, 4), 2.0), torch.tensor([2, 7]))     # line 7 does not exist: skipped
    assert c[2, :, :5].sum() == 2 * 5 * 4 * 2 and c[:2].sum() == 0
    assert U.get_layer_to_kv_cache_size_mapping_for_mixed_attn(128, 4096, [True, False, True]) == [128, 4096, 128]
    k, v = U.get_kv_shapes(256, 2, 4, 64, k_cache_transposed=True)
    assert k == (2, 4, 64, 256) and v == (2, 4, 256, 64)
    assert U.get_kv_shapes(256, 2, 4, 64, is_kv_cache_tiled=True)[1] == (2, 4, 2, 128, 64)
    # chunked prefill stitching: seq0 has 3 cached + 2 new tokens, seq1 0 cached + 3 new; block size 2
    cache = torch.arange(6 * 2, dtype=torch.float32).view(6, 2, 1, 1).expand(6, 2, 1, 4).contiguous()     # slot id in every element
    bt = torch.tensor([[4, 1, 0], [2, 0, 0]])
    slots, dc, dn = U.contexted_kv_indexing(torch.tensor([2, 3]), torch.tensor([5, 3]), bt, 2)
    assert slots.tolist() == [8, 9, 2] and dc.tolist() == [0, 1, 2] and dn.tolist() == [3, 4, 5, 6, 7]
    cur = torch.full((5, 1, 4), -1.0)
    out = U.contexted_kv(cache, cur, slots, dc, dn)
    assert out[:, 0, 0].tolist() == [8.0, 9.0, 2.0, -1.0, -1.0, -1.0, -1.0, -1.0]


def test_prefix_caching_2d_bucket_rules_and_vllm_repadding():
    """reference model_wrapper.py:923-1045 (2-D bucket choice) and :1297-1313 (vLLM re-padding)."""
    from types import SimpleNamespace
    from neuronx_distributed_inference_b200.runtime.runner import SubModelRunner
    nc = SimpleNamespace(max_context_length=1024, max_length=4096, pa_block_size=32, enable_eagle_speculation=False, allow_input_truncation=False,
                         async_mode=False, speculation_length=0, padding_side="right", pad_token_id=0)
    r = SubModelRunner.__new__(SubModelRunner)
    r.neuron_config, r.is_prefill, r.tag, r.n_active_tokens = nc, True, "context_encoding_model", 1024
    grid = [[a, p] for a in (128, 256, 512, 1024) for p in (0, 512, 1024, 2048)]
    pick = lambda a, p, **kw: r.get_target_2d_bucket_for_prefix_caching(a, p, grid, **kw)        # noqa: E731
    assert pick(100, 0) == [128, 0]
    assert pick(300, 100) == [512, 0]                         # 256 < total <= 512 corner case
    assert pick(600, 700) == [1024, 512]                      # 424 spare prefill slots absorb the prefix tail: 276 left -> 512
    assert pick(130, 2000) == [256, 2048]
    nc.enable_eagle_speculation = True
    assert pick(600, 700) == [1024, 512] and pick(900, 40) == [1024, 0] and pick(900, 100) == [1024, 512]   # +1 block; whole spare blocks only
    import pytest
    with pytest.raises(ValueError):
        pick(1020, 0)                                         # 1020 + one block does not fit
    nc.enable_eagle_speculation = False
    r.is_prefill, r.tag, r.n_active_tokens = False, "token_generation_model", 1
    tk = [[1, p] for p in (512, 1024, 2048)]
    assert r.get_target_2d_bucket_for_prefix_caching(1, 511, tk) == [1, 512]
    assert r.get_target_2d_bucket_for_prefix_caching(1, 512, tk) == [1, 1024]                    # strictly longer than the context
    assert r.get_target_2d_bucket_for_prefix_caching(1, 511, tk, strategy="second_fit") == [1, 1024]
    # vLLM padded the prompts to 300; the true longest is 140 -> re-padded to the 256 bucket
    r.is_prefill, r.seq_buckets, r.pad_token_id = True, [128, 256, 512], 0
    ids = torch.zeros(2, 300, dtype=torch.long)
    pos = torch.zeros(2, 300, dtype=torch.long)
    pos[0, :140] = torch.arange(140)
    pos[1, :90] = torch.arange(90)
    mask = (pos > 0).int()
    mask[:, 0] = 1
    i2, m2, p2, n_true = r.vllm_cte_repadding(ids, mask, pos)
    assert n_true == 140 and i2.shape == (2, 256) and m2.shape == (2, 256) and int(m2.sum()) == 230


def test_gqa_preshard_hook_matches_per_rank_sharding():
    """KV heads replicated to the TP degree: after the hook, equal dim-0 row blocks ARE the per-rank shards (weights and scales)."""
    from neuronx_distributed_inference_b200.modules.gqa import GroupQueryAttention_QKV
    from neuronx_distributed_inference_b200.parallel.state import Group
    g = Group(ranks=[0, 1, 2, 3], pg=None, rank=0)
    qkv = GroupQueryAttention_QKV(32, 8, 4, 2, tp_group=g, dtype=torch.float32)
    full_w, full_s = torch.randn((4 + 2 * 2) * 8, 32), torch.rand((4 + 2 * 2) * 8)
    sd = {"layers.0.self_attn.qkv_proj.weight": full_w.clone(), "layers.0.self_attn.qkv_proj.scale": full_s.clone()}
    assert qkv.preshard_hook(sd, "layers.0.self_attn.qkv_proj.weight")
    W, S = sd["layers.0.self_attn.qkv_proj.weight"], sd["layers.0.self_attn.qkv_proj.scale"]
    assert W.shape[0] == 4 * (1 + 2 * 1) * 8                                          # 1 q head + 1 (replicated) k + 1 v per rank
    for r in range(4):
        assert torch.equal(W.chunk(4, 0)[r], qkv._shard(full_w, r)) and torch.equal(S.chunk(4, 0)[r], qkv._shard(full_s, r))
