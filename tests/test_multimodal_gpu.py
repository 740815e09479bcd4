"""GPU tier of the image-to-text families: the CPU cases of test_multimodal_cpu.py (vision tower, image tokens scattered into the
prompt, M-RoPE decode) re-run on cuda in bf16 through the kernel path, against the same Hugging Face fp32 oracle."""
import pytest
import torch

import test_multimodal_cpu as cases

pytestmark = pytest.mark.gpu


@pytest.fixture
def on_gpu(monkeypatch):
    monkeypatch.setitem(cases.DEVICE, "on_cpu", False)
    monkeypatch.setitem(cases.DEVICE, "dtype", "bfloat16")
    monkeypatch.setitem(cases.DEVICE, "tol", 400.0)        # fp32 CPU tolerances (1e-4 .. 2e-4) -> bf16 (4e-2 .. 8e-2)
    assert torch.cuda.is_available()


def test_qwen2_vl_on_gpu(tmp_path, on_gpu):
    cases.test_qwen2_vl_matches_hf(tmp_path)


def test_pixtral_on_gpu(tmp_path, on_gpu):
    cases.test_pixtral_matches_hf(tmp_path)
