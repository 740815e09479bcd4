"""Model zoo on CPU: every family is checked against the Hugging Face implementation with random tiny weights
(prefill logits at a right-padded batch + teacher-forced decode logits).  Mirrors the reference's CPU-mode
integration tests (test_llama3_1_8b_4layer_dtype.py:249-289)."""
import pytest
import torch

from neuronx_distributed_inference_b200.config import load_pretrained_config
from neuronx_distributed_inference_b200.utils.accuracy import generate_expected_logits, teacher_forced_logits
from neuronx_distributed_inference_b200.utils.constants import get_model_cls
from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint

BASE = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
            vocab_size=160, max_position_embeddings=256)


def _hf_config(model_type):
    import transformers as T
    if model_type == "llama":
        return T.LlamaConfig(**BASE, rope_scaling=dict(rope_type="llama3", factor=8.0, high_freq_factor=4.0,
                                                       low_freq_factor=1.0, original_max_position_embeddings=64))
    if model_type == "mistral":
        return T.MistralConfig(**BASE, sliding_window=8)
    if model_type == "qwen2":
        return T.Qwen2Config(**BASE)
    if model_type == "qwen3":
        return T.Qwen3Config(**BASE, head_dim=16)
    if model_type == "gemma3":
        return T.Gemma3TextConfig(**BASE, head_dim=16, sliding_window=8, query_pre_attn_scalar=16,
                                  layer_types=["sliding_attention", "sliding_attention", "full_attention"])
    if model_type == "mixtral":
        return T.MixtralConfig(**BASE, num_local_experts=4, num_experts_per_tok=2)
    if model_type == "qwen3_moe":
        return T.Qwen3MoeConfig(**BASE, head_dim=16, num_experts=4, num_experts_per_tok=2, moe_intermediate_size=64,
                                decoder_sparse_step=1, norm_topk_prob=True, mlp_only_layers=[])
    if model_type == "dbrx":
        return T.DbrxConfig(d_model=64, n_heads=4, n_layers=2, max_seq_len=256, vocab_size=160,
                            attn_config=dict(kv_n_heads=2, clip_qkv=8.0, rope_theta=10000.0),
                            ffn_config=dict(ffn_hidden_size=64, moe_num_experts=4, moe_top_k=2, hidden_size=64))
    if model_type == "gpt_oss":
        return T.GptOssConfig(**{**BASE, "num_hidden_layers": 2}, head_dim=16, num_local_experts=4, num_experts_per_tok=2,
                              sliding_window=8, layer_types=["sliding_attention", "full_attention"])
    if model_type == "llama4":
        return T.Llama4TextConfig(**{**BASE, "num_hidden_layers": 4}, head_dim=16, num_local_experts=4, num_experts_per_tok=1,
                                  intermediate_size_mlp=128, interleave_moe_layer_step=2, attention_chunk_size=8,
                                  no_rope_layers=[1, 1, 1, 0], use_qk_norm=True, attn_temperature_tuning=True, floor_scale=4,
                                  attn_scale=0.1)
    if model_type == "deepseek":
        return T.DeepseekV3Config(**{**BASE, "num_hidden_layers": 2}, q_lora_rank=24, kv_lora_rank=16, qk_nope_head_dim=16,
                                  qk_rope_head_dim=8, v_head_dim=16, n_routed_experts=4, n_shared_experts=1,
                                  num_experts_per_tok=2, moe_intermediate_size=32, first_k_dense_replace=1, n_group=1,
                                  topk_group=1)
    raise KeyError(model_type)


FAMILIES = ["llama", "mistral", "qwen2", "qwen3", "gemma3", "mixtral", "qwen3_moe", "dbrx", "gpt_oss", "llama4", "deepseek"]


def _patch_hf_dbrx():
    """transformers 5.x evaluates the DBRX experts as ``x @ w1`` / ``h @ w2.T`` on ``[I,H]`` slices, which only
    type-checks when I == H and is the transpose of the released checkpoints' semantics (``x @ w1.T``, ``h @ w2``:
    databricks/dbrx modeling code, reference modeling_dbrx.py:51-112).  Restore the checkpoint semantics for the oracle."""
    from transformers.models.dbrx import modeling_dbrx as m

    def forward(self, x, expert_w1, expert_v1, expert_w2):
        return (self.activation_fn(x.matmul(expert_w1.t())) * x.matmul(expert_v1.t())).matmul(expert_w2)
    m.DbrxExpertGLU.forward = forward


def run_family(model_type, tmp_path, tol=2e-4, **nc_kw):
    from transformers import AutoModelForCausalLM
    if model_type == "dbrx":
        _patch_hf_dbrx()
    hf_cfg = _hf_config(model_type)
    ckpt = save_random_hf_checkpoint(hf_cfg, str(tmp_path / model_type), seed=1)
    hf = AutoModelForCausalLM.from_pretrained(ckpt, dtype=torch.float32).eval()
    cls = get_model_cls(model_type)
    nc = cls.get_neuron_config_cls()(batch_size=2, seq_len=48, max_context_length=24, torch_dtype="float32", on_cpu=True,
                                     output_logits=True, **nc_kw)
    cfg = cls.get_config_cls()(nc, load_config=load_pretrained_config(ckpt))
    app = cls(ckpt, cfg)
    app.load(None, skip_warmup=True)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(1, hf_cfg.vocab_size, (2, 14), generator=g)
    mask = torch.ones_like(ids)
    mask[1, 10:] = 0
    exp, toks = generate_expected_logits(hf, ids, mask, 12)
    got = teacher_forced_logits(app, ids, mask, toks)
    err = ((got - exp).norm() / exp.norm()).item()
    assert err < tol, f"{model_type}: relative logit error {err}"
    assert torch.equal(got.argmax(-1), exp.argmax(-1))


@pytest.mark.parametrize("model_type", FAMILIES)
def test_family_matches_hf(model_type, tmp_path):
    run_family(model_type, tmp_path)


@pytest.mark.parametrize("model_type", ["mistral", "gemma3", "gpt_oss"])
def test_rolling_sliding_window_cache(model_type, tmp_path, monkeypatch):
    """Sliding layers keep only ``window`` (8) KV slots written modulo the window; the 14-token prompts and 12 decode steps wrap
    the window three times.  Logits must still match Hugging Face (reference gpt_oss_kv_cache_manager.py:30-396)."""
    from neuronx_distributed_inference_b200.modules.kvcache.gpt_oss_kv_cache_manager import HybridKVCacheManager
    seen = []
    orig = HybridKVCacheManager.__init__

    def spy(self, *a, **kw):
        orig(self, *a, **kw)
        seen.append(self)
    monkeypatch.setattr(HybridKVCacheManager, "__init__", spy)
    run_family(model_type, tmp_path, rolling_sliding_window_cache=True)
    assert seen, "the hybrid manager was not selected"
    for m in seen:
        assert m.window == 8 and m.windowed.max_len == 8
        k, _ = m.get_kv_by_layer_id(0)                      # layer 0 slides in all three configs
        assert k.shape[2] == 8


def test_rolling_cache_rejects_prefix_features():
    from neuronx_distributed_inference_b200.config import NeuronConfig
    with pytest.raises(ValueError):
        NeuronConfig(batch_size=1, seq_len=64, rolling_sliding_window_cache=True, speculation_length=4)
