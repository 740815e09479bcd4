"""Multi-LoRA serving: per-row adapters must equal running each row through a model whose weights were merged with that
adapter (W + alpha/r * B A); LRU slot management of the dynamic mode."""
import torch

from neuronx_distributed_inference_b200.config import LoraServingConfig
from neuronx_distributed_inference_b200.modules.lora import AdapterCache
from neuronx_distributed_inference_b200.utils.testing import build_random_llama

TINY = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
            vocab_size=128, head_dim=16)


def _adapter(seed, r=4, layers=2):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    dims = dict(q_proj=(64, 64), k_proj=(32, 64), v_proj=(32, 64), o_proj=(64, 64), gate_proj=(128, 64), up_proj=(128, 64),
                down_proj=(64, 128))
    for i in range(layers):
        for p, (o, inn) in dims.items():
            mod = "self_attn" if p in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
            sd[f"base_model.model.model.layers.{i}.{mod}.{p}.lora_A.weight"] = torch.randn(r, inn, generator=g) * 0.3
            sd[f"base_model.model.model.layers.{i}.{mod}.{p}.lora_B.weight"] = torch.randn(o, r, generator=g) * 0.3
    return sd


def _merge(app, sd, alpha, r):
    """Fold the adapter into the base weights of ``app`` (fp32 CPU model)."""
    for i, layer in enumerate(app.model.layers):
        def d(p, mod):
            a = sd[f"base_model.model.model.layers.{i}.{mod}.{p}.lora_A.weight"]
            b = sd[f"base_model.model.model.layers.{i}.{mod}.{p}.lora_B.weight"]
            return (alpha / r) * b @ a
        layer.self_attn.qkv_proj.weight.data += torch.cat([d("q_proj", "self_attn"), d("k_proj", "self_attn"), d("v_proj", "self_attn")])
        layer.self_attn.o_proj.weight.data += d("o_proj", "self_attn")
        layer.mlp.gate_up_proj.weight.data += torch.cat([d("gate_proj", "mlp"), d("up_proj", "mlp")])
        layer.mlp.down_proj.weight.data += d("down_proj", "mlp")


def test_multi_lora_rows_match_merged_models():
    lc = LoraServingConfig(max_loras=3, max_lora_rank=8, target_modules=["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj",
                                                                       "up_proj", "down_proj"], lora_alpha=8)
    kw = dict(batch_size=2, seq_len=32, max_context_length=16, device="cpu", dtype="float32", seed=11, output_logits=True)
    app = build_random_llama(TINY, lora_config=lc, **kw)
    a1, a2 = _adapter(1), _adapter(2)
    app.lora_manager.add_adapter("a1", state_dict=a1, alpha=8)
    app.lora_manager.add_adapter("a2", state_dict=a2, alpha=8)
    ids = torch.randint(0, 128, (2, 7))
    aid = app.lora_manager.adapter_ids(["a2", "a1"])
    got = app(ids, attention_mask=torch.ones_like(ids), adapter_ids=aid).logits[:, -1]
    for row, sd in ((0, a2), (1, a1)):
        m = build_random_llama(TINY, **kw)
        _merge(m, sd, 8, 4)
        exp = m(ids, attention_mask=torch.ones_like(ids)).logits[row, -1]
        assert ((got[row] - exp).norm() / exp.norm()) < 1e-4
    # decode step with adapters keeps tracking the merged model
    nxt = got.argmax(-1)
    d = app(nxt.view(2, 1), position_ids=torch.full((2, 1), 7, dtype=torch.int32), adapter_ids=aid).logits[:, -1]
    m = build_random_llama(TINY, **kw)
    _merge(m, a2, 8, 4)
    m(ids, attention_mask=torch.ones_like(ids))
    e = m(nxt.view(2, 1), position_ids=torch.full((2, 1), 7, dtype=torch.int32)).logits[0, -1]
    assert ((d[0] - e).norm() / e.norm()) < 1e-4


def test_adapter_cache_lru_and_pinning():
    c = AdapterCache(2)
    assert c.allocate("a") == (0, None) and c.allocate("b") == (1, None)
    assert c.lookup("a") == 0                       # touch a -> b becomes LRU
    assert c.allocate("c") == (1, "b")
    c.pinned.add("a")
    assert c.allocate("d") == (1, "c")             # a is pinned, c goes
    c.pinned.add("d")
    try:
        c.allocate("e")
        assert False
    except RuntimeError:
        pass
    assert c.remove("a") == 0 and c.lookup("a") is None


def test_lora_serving_package_checkpoint_and_weight_manager(tmp_path):
    """reference layout modules/lora_serving/{lora_checkpoint,lora_model}.py: PEFT directory round trip, validation against the
    serving config, host pool limit, and the weight manager's tensor inventory / name -> slot resolution."""
    import json
    import pytest
    from safetensors.torch import save_file
    from neuronx_distributed_inference_b200.modules.lora_serving import LoraCheckpoint, LoraWeightManager
    from neuronx_distributed_inference_b200.modules.lora_serving.lora_module import HF_TO_FUSED, TARGETS
    assert HF_TO_FUSED["k_proj"] == ("qkv_proj", 1) and "gate_up_proj" in TARGETS
    d = tmp_path / "a1"
    d.mkdir()
    sd = _adapter(1)
    save_file(sd, str(d / "adapter_model.safetensors"))
    json.dump({"lora_alpha": 8, "r": 4}, open(d / "adapter_config.json", "w"))
    lc = LoraServingConfig(max_loras=2, max_lora_rank=8, max_cpu_loras=1, lora_alpha=8,
                           target_modules=["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"])
    ck = LoraCheckpoint(lc)
    cfg, got = ck.load_to_cpu("a1", str(d))
    assert cfg["lora_alpha"] == 8 and set(got) == set(sd) and ck.validate("a1", cfg, got) == (8, 4)
    with pytest.raises(RuntimeError):
        ck.load_to_cpu("a2", state_dict=_adapter(2))                 # host pool holds one adapter
    with pytest.raises(ValueError):
        LoraCheckpoint(LoraServingConfig(max_loras=2, max_lora_rank=2)).validate("a1", cfg, got)        # rank 4 > 2
    with pytest.raises(ValueError):
        LoraCheckpoint(LoraServingConfig(max_loras=2, max_lora_rank=8, target_modules=["q_proj"])).validate("a1", cfg, got)
    app = build_random_llama(TINY, lora_config=lc, batch_size=2, seq_len=32, max_context_length=16, device="cpu", dtype="float32", seed=11)
    wm = LoraWeightManager(lc, app.lora_manager.lm, app.lora_manager)
    n_targets = 2 * 4                                              # layers x fused projections
    assert len(wm.get_lora_tensors()) == 2 * n_targets and wm.print_lora_memory_footprint() > 0
    app.lora_manager.add_adapter("a1", state_dict=sd, alpha=8)
    assert wm.update_lora_adapter_ids(["a1", "a1"]).tolist() == app.lora_manager.adapter_ids(["a1", "a1"]).tolist()
