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


def _run(n, ckpt, port, **env_extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mp", "llama_tp_worker.py"), ckpt, "cpu"]
    env = dict(os.environ, OMP_NUM_THREADS="2", **env_extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout


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
           "--master-port", "29545", os.path.join(ROOT, "tests", "mp", "dp_sampler_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0 and '"ok": true' in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_llama_tp4_attention_dp_and_cp_match_hf(tiny_ckpt):
    # ranks replicating a KV head split the batch (decode) and the query sequence (prefill) instead of duplicating work
    _run(4, tiny_ckpt, 29546, ATTENTION_DP="2", CP="2")
