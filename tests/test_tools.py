"""Debug tooling: tensor capture, tensor replacement, KV reconstruct, snapshots, input capture, launcher."""
import os

import torch

from neuronx_distributed_inference_b200.utils.testing import build_random_llama

TINY = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
            vocab_size=128, head_dim=16)


def _app(**kw):
    return build_random_llama(TINY, batch_size=2, seq_len=32, max_context_length=16, device="cpu", dtype="float32", seed=3,
                              output_logits=True, **kw)


def test_tensor_capture_and_replacement():
    from neuronx_distributed_inference_b200.utils.tensor_capture_utils import capture_model_tensors, get_available_modules
    from neuronx_distributed_inference_b200.utils.tensor_replacement import replace_tensors
    app = _app()
    ids = torch.randint(0, 128, (2, 6))
    assert "layers.1.mlp" in get_available_modules(app)
    out, cap = capture_model_tensors(app, ["layers.0.self_attn", "layers.1.mlp"], ids, attention_mask=torch.ones_like(ids),
                                     capture_inputs=True)
    assert cap["layers.0.self_attn.outputs"].shape == (2, 16, 64) and "layers.1.mlp.inputs.0" in cap
    # replacing a module output by its own captured value is the identity; by zeros it is not
    app.reset()
    same, replaced = replace_tensors(app, {"layers.1.mlp": cap["layers.1.mlp.outputs"]}, ids, attention_mask=torch.ones_like(ids))
    assert replaced == [(0, "layers.1.mlp")] and torch.allclose(same.logits, out.logits, atol=1e-5)
    app.reset()
    diff, _ = replace_tensors(app, {"layers.1.mlp": torch.zeros_like(cap["layers.1.mlp.outputs"])}, ids,
                              attention_mask=torch.ones_like(ids))
    assert not torch.allclose(diff.logits, out.logits, atol=1e-3)


def test_kv_cache_reconstruct_matches_recomputed_keys():
    from neuronx_distributed_inference_b200.utils.kv_cache_reconstruct_utils import compare_kv_cache, reconstruct_kv_cache
    app = _app()
    ids = torch.randint(0, 128, (2, 9))
    app(ids, attention_mask=torch.ones_like(ids))
    rec = reconstruct_kv_cache(app, seq_len=9)
    assert len(rec) == 2 and rec[0][0].shape == (2, 2, 9, 16)
    assert rec[0][0].abs().sum() > 0 and all(ok for *_, ok in compare_kv_cache(rec, rec))


def test_snapshots_and_input_capture(tmp_path, monkeypatch):
    from neuronx_distributed_inference_b200.utils.debug_utils import capture_model_inputs
    from neuronx_distributed_inference_b200.utils.snapshot import maybe_register_from_env
    app = _app()
    monkeypatch.setenv("NXD_INFERENCE_CAPTURE_SNAPSHOT", "1")
    monkeypatch.setenv("NXD_INFERENCE_SNAPSHOT_OUTPUT_PATH", str(tmp_path / "snap"))
    monkeypatch.setenv("NXD_INFERENCE_SNAPSHOT_FOR_TOKENS", "2")
    assert len(maybe_register_from_env(app)) == len(app.models)
    capture_model_inputs(app, [1], str(tmp_path / "inputs"))
    ids = torch.randint(0, 128, (2, 5))
    tok = app(ids, attention_mask=torch.ones_like(ids)).tokens
    pos = torch.full((2, 1), 5, dtype=torch.int32)
    for _ in range(3):
        tok = app(tok.view(2, 1), position_ids=pos).tokens
        pos = pos + 1
    base = tmp_path / "snap"
    assert (base / "context_encoding_model" / "request0" / "rank0" / "inputs.pt").exists()
    assert (base / "token_generation_model" / "request0" / "step2" / "rank0" / "inputs.pt").exists()
    assert not (base / "token_generation_model" / "request0" / "step1").exists()
    blob = torch.load(tmp_path / "inputs" / "saved_inputs_1.pt")
    assert blob["args"][0].shape == (2, 1)


def test_launcher_command():
    from neuronx_distributed_inference_b200.scripts.nxdi_distributed_launcher import build_command
    cmd = build_command(8, script=["-m", "x"])
    assert "--nproc-per-node=8" in cmd and cmd[-2:] == ["-m", "x"]


def test_tensor_capture_sessions_on_disk_and_analysis(tmp_path):
    """Capture hook around a generation loop -> .pt files + capture_metadata.json; analysis against a reference capture pinpoints the
    first diverging module (reference tensor_capture_utils.py:22-113, 212-426)."""
    import json
    from neuronx_distributed_inference_b200.utils.tensor_capture_utils import (TensorCaptureMetadata, analyze_captured_tensors,
                                                                             get_tensor_capture_hook, list_capturable_modules_in_application)
    mods = ["layers.0.self_attn", "layers.1.mlp"]

    def run(app, d):
        hook = get_tensor_capture_hook(mods, capture_indices=[0, 2], tensor_capture_save_dir=d)
        ids = torch.randint(0, 128, (2, 6), generator=torch.Generator().manual_seed(1))
        out = hook(app, 0, ids, attention_mask=torch.ones_like(ids))
        tok = out.logits[:, -1].argmax(-1, keepdim=True)
        for step in (1, 2):
            out = hook(app, step, tok, position_ids=torch.full((2, 1), 5 + step, dtype=torch.int32))
            tok = out.logits[:, -1].argmax(-1, keepdim=True)

    good, bad = str(tmp_path / "good"), str(tmp_path / "bad")
    run(_app(), good)
    broken = _app()
    broken.model.layers[1].mlp.down_proj.weight.data.mul_(1.5)                  # the bug to find
    run(broken, bad)
    meta = json.load(open(os.path.join(good, TensorCaptureMetadata.FILE)))["tensors"]
    assert sorted(meta) == sorted(f for f in os.listdir(good) if f.endswith(".pt")) and len(meta) == 4      # steps 0 and 2 only
    assert {m["phase"] for m in meta.values()} == {"cte", "tkg"} and {m["module_name"] for m in meta.values()} == set(mods)
    rep = analyze_captured_tensors(bad, good)
    assert list(rep)[0].startswith("step0") and all(r["nan"] == 0 for r in rep.values())
    verdict = {f: r["allclose"] for f, r in rep.items()}
    assert verdict["step0_cte_layers_0_self_attn_outputs.pt"] and not verdict["step0_cte_layers_1_mlp_outputs.pt"]
    groups = list_capturable_modules_in_application(_app())
    assert "layers.0.self_attn" in groups["attention"] and "layers.1.mlp" in groups["mlp"] and "layers.0" in groups["layer"]
    assert any("norm" in n for n in groups["norm"])
