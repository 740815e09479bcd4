"""Rolling (window-sized) sliding-window KV cache on the CUDA path: same kernels as the full-length cache (fused RoPE + append +
split-KV flash decode), only the slot / horizon integers differ.  Compared against the full-length cache of the same weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
            vocab_size=512, head_dim=64, sliding_window=16)


def _mk(dtype, **kw):
    from neuronx_distributed_inference_b200.models.mistral.modeling_mistral import NeuronMistralForCausalLM
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    return build_random_llama(TINY, batch_size=2, seq_len=128, max_context_length=32, device="cuda", dtype=dtype, seed=7,
                              app_cls=NeuronMistralForCausalLM, output_logits=True, **kw)


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-4), ("bfloat16", 6e-2)])
def test_rolling_cache_matches_full_cache_gpu(dtype, tol):
    from neuronx_distributed_inference_b200.modules.kvcache.gpt_oss_kv_cache_manager import HybridKVCacheManager
    full, roll = _mk(dtype), _mk(dtype, rolling_sliding_window_cache=True)
    assert isinstance(roll.model.kv_mgr, HybridKVCacheManager) and roll.model.kv_mgr.bytes() < full.model.kv_mgr.bytes()
    torch.manual_seed(1)
    ids = torch.randint(1, 512, (2, 27))               # prompt already longer than the window
    mask = torch.ones_like(ids)
    mask[1, 20:] = 0
    a, b = full(ids, attention_mask=mask), roll(ids, attention_mask=mask)
    la, lb = a.logits.float(), b.logits.float()
    assert ((la - lb).norm() / la.norm()).item() < tol
    pos = mask.sum(1, keepdim=True).int()
    tok = a.tokens.view(2, 1).cpu()
    for i in range(40):                                 # wraps the 16-slot window twice more
        a, b = full(tok, position_ids=pos + i), roll(tok, position_ids=pos + i)
        la, lb = a.logits.float(), b.logits.float()
        assert ((la - lb).norm() / la.norm()).item() < tol, f"step {i}"
        tok = a.tokens.view(2, 1).cpu()
