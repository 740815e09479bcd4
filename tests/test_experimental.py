"""Experimental functional API: functional Llama-3 == the application model on the same weights; YAML per-tag configs; bucketing."""
import torch

from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter
from neuronx_distributed_inference_b200.utils.testing import build_random_llama


def test_functional_llama3_matches_application():
    from neuronx_distributed_inference_b200.experimental.core import generate
    from neuronx_distributed_inference_b200.experimental.models.llama3.model import Llama3, Llama3Args
    tiny = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                vocab_size=128, head_dim=16, rope_theta=10000.0)
    app = build_random_llama(tiny, batch_size=2, seq_len=48, max_context_length=16, device="cpu", dtype="float32", seed=4)
    ids = torch.randint(1, 128, (2, 7))
    mask = torch.ones_like(ids)
    mask[1, 5:] = 0
    ref = HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=6)
    args = Llama3Args(dim=64, n_layers=2, n_heads=4, n_kv_heads=2, vocab_size=128, ffn_dim=128, norm_eps=1e-5, rope_theta=10000.0,
                      max_batch_size=2, max_seq_len=48, dtype=torch.float32)
    model = Llama3(args, dict(app.model.state_dict()), torch.device("cpu"))
    got = generate(model, ids, mask, max_new_tokens=6)
    for b in range(2):
        n = int(mask[b].sum()) + 6
        assert got[b, :n].tolist() == ref[b, :n].tolist()


def test_yaml_config_handler_and_bucketing():
    from neuronx_distributed_inference_b200.experimental.core import BucketingProcessor, NeuronConfigHandler, load_yaml_config
    h = NeuronConfigHandler(load_yaml_config("""
common: {batch_size: 2, seq_len: 256, torch_dtype: bfloat16}
context_encoding_model: {buckets: [64, 128]}
token_generation_model: {buckets: [128, 256], cuda_graphs: true}
"""))
    assert h.tags() == ["context_encoding_model", "token_generation_model"]
    assert h.for_tag("context_encoding_model").buckets == [64, 128] and h.for_tag("token_generation_model").seq_len == 256
    p = BucketingProcessor([8, 16])
    t, m, b = p(torch.ones(2, 11, dtype=torch.long))
    assert b == 16 and t.shape == (2, 16) and int(m.sum()) == 22


def test_functional_ops_on_cpu():
    from neuronx_distributed_inference_b200.experimental import functional as F
    x = torch.randn(2, 3, 32)
    wg, wu, wd = torch.randn(64, 32) * 0.1, torch.randn(64, 32) * 0.1, torch.randn(32, 64) * 0.1
    ref = (torch.nn.functional.silu(x @ wg.T) * (x @ wu.T)) @ wd.T
    assert torch.allclose(F.gated_mlp(x, wg, wu, wd), ref, atol=1e-5)
    q, k, v = F.qkv_proj(x, torch.randn(4 * 8 + 2 * 2 * 8, 32), 4, 2, 8)
    assert q.shape == (2, 3, 4, 8) and k.shape == (2, 3, 2, 8)
    assert F.causal_scaled_dot_product_attention(q, k, v).shape == (2, 3, 4, 8)
