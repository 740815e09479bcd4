"""inference_demo end to end on CPU: tiny random Llama checkpoint + a word-level tokenizer built offline ->
compile, load, token-matching accuracy check against the HF model, generation, benchmark report, input capture, snapshots."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tiny_model_dir(tmp_path_factory):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import LlamaConfig, PreTrainedTokenizerFast
    from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint
    d = str(tmp_path_factory.mktemp("cli_model"))
    words = ["<pad>", "<s>", "</s>", "<unk>"] + [f"w{i}" for i in range(160)] + "i believe the meaning of life is color sky hello world".split()
    vocab = {w: i for i, w in enumerate(dict.fromkeys(words))}
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, pad_token="<pad>", bos_token="<s>", eos_token="</s>", unk_token="<unk>")
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      vocab_size=len(vocab), max_position_embeddings=128, pad_token_id=0, bos_token_id=1, eos_token_id=2)
    save_random_hf_checkpoint(cfg, d, seed=0)
    fast.save_pretrained(d)
    return d


def test_inference_demo_runs_end_to_end_on_cpu(tiny_model_dir, tmp_path):
    art = str(tmp_path / "artifacts")
    rep = str(tmp_path / "benchmark_report.json")
    cap = str(tmp_path / "captured")
    env = dict(os.environ, PYTHONPATH=ROOT, NXD_INFERENCE_CAPTURE_SNAPSHOT="1", NXD_INFERENCE_SNAPSHOT_OUTPUT_PATH=str(tmp_path / "snap"),
               OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "neuronx_distributed_inference_b200.inference_demo", "--model-type", "llama", "--task-type", "causal-lm", "run",
           "--model-path", tiny_model_dir, "--compiled-model-path", art, "--torch-dtype", "float32", "--tp-degree", "1", "--batch-size", "2",
           "--max-context-length", "16", "--seq-len", "32", "--on-device-sampling", "--on-cpu", "--pad-token-id", "0",
           "--prompt", "i believe the meaning of life is", "--prompt", "the color of the sky is",
           "--check-accuracy-mode", "token-matching", "--num-tokens-to-check", "8", "--benchmark", "--num-runs", "2",
           "--benchmark-report-path", rep, "--capture-indices", "1", "--input-capture-save-dir", cap]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "Accuracy check passed" in r.stdout and "Benchmark completed" in r.stdout
    report = json.load(open(rep))
    assert "e2e_model" in report and report["e2e_model"]["latency_ms_p50"] > 0 and "throughput" in report["e2e_model"]
    assert os.path.exists(os.path.join(art, "neuron_config.json")) or any(f.endswith(".json") for f in os.listdir(art))
    assert os.path.exists(os.path.join(cap, "saved_inputs_1.pt"))
    assert os.path.isdir(os.path.join(str(tmp_path / "snap"), "context_encoding_model"))
