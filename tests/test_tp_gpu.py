"""Tensor parallel on real GPUs: NCCL bootstrap + the fused GEMV->all-reduce kernel over NVLink peer memory."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _ckpt(tmp):
    from transformers import LlamaConfig
    from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint
    cfg = LlamaConfig(hidden_size=1024, intermediate_size=2048, num_hidden_layers=2, num_attention_heads=8,
                      num_key_value_heads=2, head_dim=128, vocab_size=2048, max_position_embeddings=256,
                      tie_word_embeddings=False)
    return save_random_hf_checkpoint(cfg, tmp)


@pytest.mark.parametrize("n", [2, 4, 8])
def test_llama_tp_fused_allreduce_matches_hf(n, tmp_path):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    ckpt = _ckpt(str(tmp_path / "ckpt"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + n), os.path.join(ROOT, "tests", "mp", "llama_tp_worker.py"), ckpt, "cuda"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"ok": true' in r.stdout


@pytest.mark.parametrize("n", [2, 8])
def test_symmetric_heap_nvls_collectives_and_fused_gemm(n):
    """VMM symmetric heap + NVLS multicast: in-switch all-reduce / reduce-scatter / all-gather and the fused tcgen05 GEMM ->
    reduce-scatter / all-reduce kernels vs NCCL (tests/mp/nvls_worker.py asserts the numerics on every rank)."""
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + n), os.path.join(ROOT, "tests", "mp", "nvls_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"ok": true' in r.stdout


@pytest.mark.parametrize("n", [2, 8])
def test_fused_gemv_allreduce_kernel_bf16_and_8bit_weights(n):
    """The one-kernel GEMV -> all-reduce -> +residual (LL protocol over peer memory) for bf16, fp8 and int8 weights: numerics vs
    GEMM + NCCL all-reduce, bitwise equality across ranks, CUDA-graph replay with fresh tags (tests/mp/symm_worker.py)."""
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + n), os.path.join(ROOT, "tests", "mp", "symm_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "identical_across_ranks=False" not in r.stdout and "quantised torch.int8" in r.stdout
