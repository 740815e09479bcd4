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


def test_cli_accepts_every_reference_flag_spelling(tmp_path):
    """The flags of the reference's `inference_demo run` (inference_demo.py:99-408) that are aliases or feed nested configs."""
    import json
    from neuronx_distributed_inference_b200 import inference_demo as demo
    from neuronx_distributed_inference_b200.models.llama.modeling_llama import NeuronLlamaForCausalLM
    from neuronx_distributed_inference_b200.models.mixtral.modeling_mixtral import NeuronMixtralForCausalLM
    lj = tmp_path / "lora.json"
    lj.write_text(json.dumps({"lora-ckpt-dir": "/adapters", "lora-ckpt-paths": {"a": "a.pt"}, "lora-ckpt-paths-cpu": {"b": "b.pt", "c": "c.pt"}}))
    base = ["--model-type", "llama", "--task-type", "causal-lm", "run", "--model-path", "m", "--compiled-model-path", "c", "--prompt", "hi",
            "--on-cpu", "--batch-size", "1", "--seq-len", "64", "--max-context-length", "32"]
    a = demo.parse_args(base + ["--enable-block-kv-layout", "--pa-num-blocks", "8", "--pa-block-size", "16", "--enable-prefix-caching",
                                "--kv-cache-quant", "--no-kv-direct-cast", "--k-quant-method", "per_key_symmetric",
                                "--v-quant-method", "per_key_symmetric", "--kv-quant-dtype", "float8_e4m3fn", "--cast-type", "as-declared",
                                "--start_rank_id", "0", "--local_ranks_size", "1", "--enable-lora", "--enable-dynamic-multi-lora",
                                "--lora-ckpt-json", str(lj), "--qkv-nki-kernel-enabled", "--qkv-cte-nki-kernel-fuse-rope",
                                "--attn-block-tkg-nki-kernel-cache-update", "--attn-block-tkg-nki-kernel-cascaded-attention",
                                "--strided-context-parallel-kernel-enabled", "--enable-output-completion-notifications",
                                "--logical-neuron-cores", "2", "--enable-torch-dist"])
    nc = demo.create_neuron_config(NeuronLlamaForCausalLM, a)
    assert nc.is_block_kv_layout and nc.is_prefix_caching and nc.cast_type == "as-declared"
    assert nc.kv_cache_quant and nc.kv_quant_config.scale_mode == "per_head" and nc.kv_quant_config.dtype == "float8_e4m3fn"
    assert nc.lora_config.dynamic_multi_lora and nc.lora_config.lora_ckpt_paths == {"a": "/adapters/a.pt"}
    assert set(nc.lora_config.lora_ckpt_paths_cpu) == {"b", "c"} and a.enable_torch_dist
    b = demo.parse_args([x if x != "llama" else "mixtral" for x in base] + ["--router-act-fn", "sigmoid", "--router-dtype", "bfloat16",
                                                                          "--enable-chunked-prefill", "--pa-num-blocks", "8",
                                                                          "--pa-block-size", "16"])
    mc = demo.create_neuron_config(NeuronMixtralForCausalLM, b)
    assert mc.router_config == {"act_fn": "sigmoid", "dtype": "bfloat16"} and mc.is_chunked_prefill and mc.is_block_kv_layout
