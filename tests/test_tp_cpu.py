"""Tensor-parallel correctness on CPU: real multi-process gloo groups (world_size 2), checked against HF."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tiny_ckpt(tmp_path_factory):
    from transformers import LlamaConfig
    from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=320, max_position_embeddings=128, tie_word_embeddings=False)
    return save_random_hf_checkpoint(cfg, str(tmp_path_factory.mktemp("ckpt")))


def _free_port(hint: int = 0) -> int:
    """A currently unused TCP port on 127.0.0.1 (the fixed numbers in the tests are only hints: another job on the machine may own them)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(n, ckpt, port, **env_extra):
    port = _free_port(port)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp", "llama_tp_worker.py"), ckpt, "cpu"]
    env = dict(os.environ, OMP_NUM_THREADS="2", **env_extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout
    return r.stdout


def test_llama_tp2_gloo_matches_hf(tiny_ckpt):
    _run(2, tiny_ckpt, 29541)


def test_llama_tp4_gloo_kv_replication_matches_hf(tiny_ckpt):
    # 2 KV heads at TP=4 -> REPLICATE_TO_TP_DEGREE path of the GQA plan
    _run(4, tiny_ckpt, 29542)


def test_llama_tp4_flash_decoding_matches_hf(tiny_ckpt):
    # KV heads replicated on 2 ranks each -> the sequence is sharded inside each pair (flash decoding)
    _run(4, tiny_ckpt, 29543, FLASH_DECODING="1")


def test_llama_tp2_sequence_parallel_matches_hf(tiny_ckpt):
    # residual stream sharded along the sequence during prefill (all-gather before column-, reduce-scatter after row-parallel)
    _run(2, tiny_ckpt, 29544, SEQUENCE_PARALLEL="1")


def test_data_parallel_sampler_matches_distributed_sampler():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port(29545)), os.path.join(ROOT, "tests", "mp", "dp_sampler_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0 and '"ok": true' in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_llama_tp4_attention_dp_and_cp_match_hf(tiny_ckpt):
    # ranks replicating a KV head split the batch (decode) and the query sequence (prefill) instead of duplicating work
    _run(4, tiny_ckpt, 29546, ATTENTION_DP="2", CP="2")


def _contrib_cfg(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("_contrib_cfgs", os.path.join(ROOT, "tests", "test_contrib_cpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod._cfg(name)


def test_contrib_families_tp2_gloo_match_hf(tmp_path):
    """The sharding metadata of the contrib blocks (3-way fused conv projection, head-sharded RG-LRU gates, replicated Mamba-2 mixer,
    q-head-aligned attention gate, expert sharding, per-head LayerNorm) non-gated experts) under a real 2-rank gloo group — eleven configurations in one launch."""
    from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint
    # falcon_h1: Mamba-2 heads / groups sharded (2 groups on 2 ranks); falcon_h1_one_group: ONE group spanning both ranks with the gated
    # group norm (its sum of squares is combined across the ranks)
    names = ["lfm2", "recurrent_gemma", "falcon_h1", "falcon_h1_one_group", "afmoe", "phimoe", "persimmon", "nemotron_h", "falcon_mamba", "bloom", "qwen3_next"]
    ckpts = [save_random_hf_checkpoint(_contrib_cfg(n), str(tmp_path / n), seed=4) for n in names]
    out = _run(2, ",".join(ckpts), 29560, MODEL_TYPE=",".join(n.replace("_one_group", "") for n in names), DUMP_AFTER="500")
    assert out.count('"ok": true') == len(names)


def test_gemma3_tp2_rolling_sliding_window_cache(tmp_path):
    """Mixed sliding / global layers with the window-sized rolling cache, sharded over 2 ranks: prompts (12 tokens) and decode steps
    wrap the 8-slot window."""
    import transformers as T
    from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint
    cfg = T.Gemma3TextConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
                             vocab_size=160, max_position_embeddings=256, head_dim=16, sliding_window=8, query_pre_attn_scalar=16,
                             layer_types=["sliding_attention", "sliding_attention", "full_attention"])
    ckpt = save_random_hf_checkpoint(cfg, str(tmp_path / "g3"), seed=4)
    _run(2, ckpt, 29570, MODEL_TYPE="gemma3", ROLLING_SWA="1")


def test_weight_gathered_matmul_gloo():
    """EAGLE weight-gather projections (reference eagle/utils.py:65-205): y = x @ all_gather(W)^T tiled over K / looped over N."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port(29575)), os.path.join(ROOT, "tests", "mp", "weight_gather_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout


@pytest.mark.parametrize("i,mode", [(0, "cp"), (1, "cfg")])
def test_flux_dp2_context_and_cfg_parallel(i, mode):
    """FLUX with world = 2 x tp (reference application.py:33-65): image tokens split over the two replicas (context parallel), or the
    conditional / unconditional branches of true CFG split (CFG parallel) — both equal the single-replica result."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port(29580 + i)), os.path.join(ROOT, "tests", "mp", "flux_dp2_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout


def test_llama_general_cp_and_attention_dp_meshes_match_hf(tiny_ckpt):
    # degrees that are NOT the KV replication factor: TP2 (one kv head per rank, no replication) with cp = dp = 2, and TP4
    # (replication factor 2) with cp = 4 — blocks of adjacent ranks that gather their K/V heads
    _run(2, tiny_ckpt, 29561, ATTENTION_DP="2", CP="2")
    _run(4, tiny_ckpt, 29562, ATTENTION_DP="2", CP="4")
    _run(2, tiny_ckpt, 29563, ATTENTION_DP="2")            # general DP alone: the prefill is ordinary attention, the cache write gathers heads
    _run(2, tiny_ckpt, 29564, CP="2", STRIDED_CP="1")      # strided sequence split (causal load balance)
