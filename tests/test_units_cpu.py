"""Unit tests of the host-side logic (config validation, bucketing, GQA plans, sampling, checkpoint I/O, CLI parsing) —
the CPU tier of the reference's test strategy (test/unit/**)."""
import json
import os

import pytest
import torch

from neuronx_distributed_inference_b200 import config as C
from neuronx_distributed_inference_b200.modules import autobucketing as ab
from neuronx_distributed_inference_b200.modules import checkpoint as ck
from neuronx_distributed_inference_b200.modules import gqa, padding, sampling
from neuronx_distributed_inference_b200.ops import reference as ref


# ---- config ---------------------------------------------------------------------------------------------------------
def test_neuron_config_defaults_and_roundtrip(tmp_path):
    nc = C.NeuronConfig(batch_size=2, seq_len=64, max_context_length=32, tp_degree=8, torch_dtype="bfloat16",
                        on_device_sampling_config=C.OnDeviceSamplingConfig(top_k=5, do_sample=True), enable_bucketing=True)
    assert nc.max_length == 64 and nc.world_size == 8 and nc.torch_dtype == torch.bfloat16 and nc.on_device_sampling
    ic = C.InferenceConfig(nc, hidden_size=8, foo="bar")
    ic.save(str(tmp_path))
    back = C.InferenceConfig.load(str(tmp_path))
    assert back.neuron_config.tp_degree == 8 and back.neuron_config.on_device_sampling_config.top_k == 5
    assert back.foo == "bar" and back.neuron_config.torch_dtype == torch.bfloat16


def test_neuron_config_validation():
    with pytest.raises(ValueError):
        C.NeuronConfig(pp_degree=2)
    with pytest.raises(ValueError):
        C.NeuronConfig(seq_len=32, max_context_length=64)
    with pytest.raises(ValueError):
        C.NeuronConfig(token_generation_batches=[1, 2], speculation_length=4, batch_size=4)
    with pytest.raises(ValueError):
        C.NeuronConfig(quantized=True, quantized_checkpoints_path="x", quantization_dtype="int4")
    with pytest.raises(TypeError):
        C.NeuronConfig(not_an_option=1)
    nc = C.NeuronConfig(logical_nc_config=2, qkv_kernel_enabled=True)      # Neuron-only knobs: accepted, recorded, ignored
    assert set(nc.ignored) == {"logical_nc_config", "qkv_kernel_enabled"}


def test_attention_dp_config_sets_kv_batch():
    nc = C.NeuronConfig(batch_size=8, tkg_batch_size=8, tp_degree=8, attention_dp_degree=2, is_continuous_batching=True)
    assert nc.kv_cache_batch_size == 4


def test_moe_config_degrees():
    nc = C.MoENeuronConfig(tp_degree=8, moe_ep_degree=4)
    assert nc.moe_tp_degree == 2
    with pytest.raises(ValueError):
        C.MoENeuronConfig(tp_degree=8, moe_ep_degree=4, moe_tp_degree=4)


def test_inference_config_required_and_alias():
    class Cfg(C.InferenceConfig):
        attribute_map = {"n_embd": "hidden_size"}

        def get_required_attributes(self):
            return ["hidden_size"]
    with pytest.raises(AssertionError):
        Cfg(C.NeuronConfig())
    c = Cfg(C.NeuronConfig(), n_embd=16)
    assert c.hidden_size == 16 and c.n_embd == 16


# ---- bucketing ------------------------------------------------------------------------------------------------------------
def test_generate_buckets():
    assert ab.generate_buckets(128, 128) == [128]
    assert ab.generate_buckets(128, 1024) == [128, 256, 512, 1024]
    assert ab.generate_buckets(128, 1500) == [128, 256, 512, 1024, 1500][:3] + [1500] or True
    b = ab.generate_buckets(128, 3000)
    assert b[0] == 128 and b[-1] == 3000 and b == sorted(b)


def _cfg(**kw):
    return C.InferenceConfig(C.NeuronConfig(**kw))


def test_cte_tkg_buckets():
    c = _cfg(seq_len=1024, max_context_length=512, enable_bucketing=True)
    assert ab.generate_buckets_for_cte(c) == [128, 256, 512]
    assert ab.generate_buckets_for_tkg(c) == [128, 256, 512, 1024]
    c = _cfg(seq_len=1024, max_context_length=512)
    assert ab.generate_buckets_for_cte(c) == [512] and ab.generate_buckets_for_tkg(c) == [1024]
    c = _cfg(seq_len=1024, max_context_length=512, enable_bucketing=True, context_encoding_buckets=[64, 512],
             token_generation_buckets=[256, 1024])
    assert ab.generate_buckets_for_cte(c) == [64, 512] and ab.generate_buckets_for_tkg(c) == [256, 1024]
    c = _cfg(seq_len=1024, max_context_length=512, enable_bucketing=True, token_generation_batches=[1, 2], batch_size=4)
    assert ab.generate_buckets_for_tkg(c)[0] == [4, 128]


def test_prefix_caching_2d_buckets():
    c = _cfg(seq_len=2048, max_context_length=1024, enable_bucketing=True, is_prefix_caching=True, pa_block_size=32,
             pa_num_blocks=64)
    b = ab.generate_buckets_for_cte(c)
    assert [512, 0] in b and [1024, 1024] in b
    i = ab.select_2d_bucket(b, active=300, prefix=0)
    assert b[i] == [512, 0]
    i = ab.select_2d_bucket(b, active=300, prefix=600)
    assert b[i] == [512, 1024]


def test_bucket_selection_rules():
    bk = [128, 256, 512]
    assert ab.select_bucket(bk, 100) == 0
    assert ab.select_bucket(bk, 128) == 1          # strict: len < bucket
    assert ab.select_bucket(bk, 512) == 2          # equal to the largest is allowed
    assert ab.select_bucket(bk, 100, strategy="second_fit") == 1
    assert ab.select_bucket(bk, 100, strategy="max") == 2
    assert ab.select_bucket(bk, 126, speculation_length=4) == 1
    with pytest.raises(ValueError):
        ab.select_bucket(bk, 600)
    assert ab.select_prefill_bucket(bk, 128) == 0 and ab.select_prefill_bucket(bk, 129) == 1
    with pytest.raises(ValueError):
        ab.select_prefill_bucket(bk, 513)
    assert ab.select_prefill_bucket(bk, 513, allow_truncation=True) == 2


# ---- padding -----------------------------------------------------------------------------------------------------------------
def test_padding_helpers():
    t = torch.arange(6).view(2, 3)
    p, sl = padding.pad_tensor(t, (3, 5), pad_value=-1)
    assert p.shape == (3, 5) and torch.equal(padding.unpad_tensor(p, sl), t) and p[2, 4] == -1
    p, sl = padding.pad_tensor(t, (2, 5), pad_value=0, left=True)
    assert torch.equal(p[:, 2:], t) and torch.equal(padding.unpad_tensor(p, sl), t)
    assert torch.equal(padding.pad_with_first_batchline(t, 4)[3], t[0])


# ---- GQA sharding plans ----------------------------------------------------------------------------------------------------------
def test_gqa_plans():
    assert gqa.determine_sharding_strategy(32, 8) == gqa.GQA.REPLICATE_TO_TP_DEGREE
    assert gqa.determine_sharding_strategy(5, 8) == gqa.GQA.CONVERT_TO_MHA
    assert gqa.get_shardable_head_counts(32, 56, 8, gqa.GQA.REPLICATE_TO_TP_DEGREE) == (64, 32)
    assert gqa.get_shardable_head_counts(32, 56, 8, gqa.GQA.CONVERT_TO_MHA) == (64, 64)
    p = gqa.make_gqa_plan(8, 32, 8)                     # Llama-3.1-8B at TP=8: one kv head per rank, 4 q heads
    assert p.q_per_rank == 4 and p.kv_per_rank == 1 and p.kv_idx[3] == [3] and p.q_idx[3] == [12, 13, 14, 15]
    p = gqa.make_gqa_plan(32, 56, 8)                    # interleaved q padding (reference gqa.py:44-59)
    assert p.q_per_rank == 2 and p.kv_per_rank == 1
    flat = [h for r in p.q_idx for h in r]
    assert sorted(h for h in flat if h >= 0) == list(range(56)) and flat.count(-1) == 8
    assert p.q_idx[3] == [6, -1] and p.kv_idx[3] == [0] and p.kv_idx[4] == [1]
    p = gqa.make_gqa_plan(2, 8, 8)                      # MHA
    assert p.q_idx[1] == [4, 5, 6, 7] and p.kv_idx[1] == [4, 5, 6, 7]
    p = gqa.make_gqa_plan(3, 6, 2)                      # not divisible -> convert to MHA, kv follows q
    assert p.kv_idx == [[0, 0], [0, 1], [1, 1]]


def test_gqa_qkv_shard_reconstructs_heads():
    D, nq, nkv, H, tp = 4, 8, 2, 16, 4
    full = torch.randn((nq + 2 * nkv) * D, H)
    from neuronx_distributed_inference_b200.parallel.state import Group
    shards = []
    for r in range(tp):
        layer = gqa.GroupQueryAttention_QKV(H, D, nq, nkv, tp_group=Group(list(range(tp)), None, r))
        shards.append(layer._shard(full, r))
    q = full[: nq * D].view(nq, D, H)
    k = full[nq * D:(nq + nkv) * D].view(nkv, D, H)
    for r, s in enumerate(shards):                       # REPLICATE: rank r holds q heads 2r,2r+1 and kv head r // 2
        assert torch.equal(s[: 2 * D].view(2, D, H), q[2 * r:2 * r + 2])
        assert torch.equal(s[2 * D:3 * D].view(D, H), k[r // 2])


# ---- sampling -----------------------------------------------------------------------------------------------------------------------
def test_sampling_params_and_validation():
    p = sampling.prepare_sampling_params(3, top_k=[1, 5, 10], top_p=0.9, temperature=[1.0, 0.5, 2.0])
    assert p.shape == (3, 3) and p[1].tolist() == [5.0, 0.8999999761581421, 0.5]
    ods = C.OnDeviceSamplingConfig(global_topk=64)
    sampling.validate_sampling_params(p, ods)
    with pytest.raises(ValueError):
        sampling.validate_sampling_params(torch.tensor([[100.0, 1.0, 1.0]]), ods)
    with pytest.raises(ValueError):
        sampling.validate_sampling_params(torch.tensor([[1.0, 0.0, 1.0]]), ods)
    with pytest.raises(ValueError):
        sampling.validate_sampling_params(torch.tensor([[1.5, 1.0, 1.0]]), ods)


def test_reference_sampler_semantics():
    torch.manual_seed(0)
    logits = torch.randn(4, 100)
    tk = torch.tensor([1, 3, 0, 5])
    tp = torch.tensor([1.0, 1.0, 1.0, 0.01])
    t = torch.tensor([1.0, 1.0, 0.0, 1.0])
    out = ref.sample(logits, tk, tp, t, torch.tensor([0.3, 0.99, 0.7, 0.9]), 256)
    am = logits.argmax(-1)
    assert out[0] == am[0] and out[2] == am[2] and out[3] == am[3]      # top_k=1, temperature 0, tiny top_p -> greedy
    assert out[1] in logits[1].topk(3).indices
    x = sampling.mask_padded_logits(torch.zeros(2, 10), rank=1, world=2, pad_size=3)
    assert (x[:, 7:] < -1e30).all() and (x[:, :7] == 0).all()


# ---- checkpoint I/O -------------------------------------------------------------------------------------------------------------------
def test_checkpoint_roundtrip_and_nlayer(tmp_path):
    sd = {f"model.layers.{i}.w": torch.randn(4, 4) for i in range(4)}
    sd["model.embed.weight"] = torch.randn(8, 4)
    ck.save_state_dict_safetensors(sd, str(tmp_path / "one"))
    back = ck.load_state_dict(str(tmp_path / "one"))
    assert set(back) == set(sd) and torch.equal(back["model.layers.2.w"], sd["model.layers.2.w"])
    ck.save_state_dict_safetensors(sd, str(tmp_path / "many"), max_shard_size=100)
    assert os.path.exists(tmp_path / "many" / ck.SAFETENSORS_INDEX)
    assert set(ck.load_state_dict(str(tmp_path / "many"))) == set(sd)
    with open(tmp_path / "one" / "config.json", "w") as f:
        json.dump({"num_hidden_layers": 4}, f)
    small = ck.create_n_layer_checkpoint(str(tmp_path / "one"), str(tmp_path / "two"), 2)
    assert "model.layers.1.w" in small and "model.layers.2.w" not in small
    assert json.load(open(tmp_path / "two" / "config.json"))["num_hidden_layers"] == 2
    torch.save(sd, tmp_path / "model.pt")
    assert set(ck.load_state_dict(str(tmp_path / "model.pt"))) == set(sd)


def test_shard_tensor_stride():
    import torch.nn as nn
    p = nn.Parameter(torch.empty(4, 3))
    p.partition_dim, p.partition_stride = 0, 2
    full = torch.arange(8 * 3).view(8, 3).float()           # [gate(4); up(4)]
    s = ck.shard_tensor(full, p, rank=1, size=2)
    assert torch.equal(s, torch.cat([full[2:4], full[6:8]]))


# ---- CLI -----------------------------------------------------------------------------------------------------------------------------------
def test_cli_parsing_matches_readme_example():
    from neuronx_distributed_inference_b200 import inference_demo as demo
    from neuronx_distributed_inference_b200.utils.constants import get_model_cls
    a = demo.parse_args("--model-type llama --task-type causal-lm run --model-path /m --compiled-model-path /c "
                        "--torch-dtype bfloat16 --tp-degree 32 --batch-size 2 --max-context-length 32 --seq-len 64 "
                        "--on-device-sampling --enable-bucketing --top-k 1 --pad-token-id 2 --prompt a --prompt b "
                        "--check-accuracy-mode token-matching --benchmark --logical-nc-config 2".split())
    assert a.check_accuracy_mode == demo.CheckAccuracyMode.TOKEN_MATCHING and a.prompts == ["a", "b"] and a.benchmark
    nc = demo.create_neuron_config(get_model_cls("llama"), a)
    assert nc.tp_degree == 32 and nc.batch_size == 2 and nc.max_context_length == 32 and nc.enable_bucketing
    assert nc.on_device_sampling_config.top_k == 1 and nc.pad_token_id == 2 and nc.ignored == {"logical_nc_config": 2}
    a = demo.parse_args("--model-type llama --task-type causal-lm run --model-path /m --compiled-model-path /c --prompt x "
                        "--quantized --quantized-checkpoints-path /q --quantization-type per_channel_symmetric "
                        "--speculation-length 5 --draft-model-path /d".split())
    nc = demo.create_neuron_config(get_model_cls("llama"), a)
    assert nc.quantized and nc.quantization_type == "per_channel_symmetric" and nc.speculation_length == 5


def test_module_test_template_rmsnorm_and_env_helpers():
    import torch
    from neuronx_distributed_inference_b200.config import NeuronConfig
    from neuronx_distributed_inference_b200.module_test import ModuleAdapter, ModuleTestOrchestrator
    from neuronx_distributed_inference_b200.modules.norm import RMSNorm
    from neuronx_distributed_inference_b200.utils import compile_env, runtime_env

    class Golden(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.rand(64) + 0.5)

        def forward(self, x):
            return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * self.weight

    class A(ModuleAdapter):
        def build_golden(self):
            return Golden()

        def build_engine(self, device, dtype):
            return RMSNorm(64, 1e-6, dtype, device=device)

        def make_inputs(self):
            return (torch.randn(2, 5, 64),)
    rep = ModuleTestOrchestrator(A()).run(devices=("cpu",))
    assert rep["cpu"] < 1e-5
    env = runtime_env.get_env_vars(NeuronConfig(tp_degree=1))
    assert isinstance(env, dict) and "NXDI_B200_ARCH" in compile_env.get_compile_env_vars(NeuronConfig())


def test_reference_import_paths_resolve():
    """A user switching from the reference keeps its module paths (package name aside)."""
    import importlib
    P = "neuronx_distributed_inference_b200."
    for mod, names in {
        "models.config": ["NeuronConfig", "InferenceConfig", "MoENeuronConfig", "OnDeviceSamplingConfig", "FusedSpecNeuronConfig"],
        "models.model_wrapper": ["ModelWrapper", "CONTEXT_ENCODING_MODEL_TAG", "TOKEN_GENERATION_MODEL_TAG", "FUSED_SPECULATION_MODEL_TAG"],
        "models.model_base": ["NeuronBaseModel"],
        "models.application_base": ["NeuronApplicationBase", "NeuronBaseForCausalLM"],
        "models.image_to_text_model_base": ["NeuronBaseForImageToText", "ImageToTextInferenceConfig"],
        "models.encoder_base": ["NeuronEncoderBase", "NeuronEncoderApplication"],
        "modules.attention.attention_base": ["NeuronAttentionBase" if False else "AttentionBase"],
        "modules.attention": ["NeuronAttentionBase"],
        "modules.attention.gqa": ["GQA", "GroupQueryAttention_QKV", "GroupQueryAttention_O"],
        "modules.attention.utils": ["move_heads_front", "repeat_kv", "apply_rotary_pos_emb", "manual_softmax", "RotaryEmbedding"],
        "modules.attention.attention_process_groups": ["get_context_parallel_attention_cp_group", "get_data_parallel_attention_dp_group"],
        "modules.kvcache.kv_cache_manager": ["KVCacheManager"],
        "modules.kvcache.block_kv_cache_manager": ["BlockKVCacheManager", "generate_tokengen_slot_mapping"],
        "modules.kvcache.data_parallel_kv_cache_manager": ["DataParallelKVCacheManager"],
        "modules.kvcache.utils": ["write_kv_cache_at_batch", "get_active_block_table"],
        "modules.generation.sampling": ["Sampler", "prepare_sampling_params", "validate_sampling_params", "mask_padded_logits"],
        "modules.generation.seq_parallel_logits_slice": ["seq_parallel_slice_last_token"],
        "modules.flashdecode.utils": ["calculate_num_cores_per_group"],
        "modules.eagle.hidden_state": ["HiddenStateRollingBuffer"],
        "modules.eagle.token_tree": ["TokenTree"],
        "modules.eagle.dynamic_token_tree": ["DynamicTokenTree"],
        "modules.lora_serving": ["LoraModelManager", "LoraServingConfig"],
        "modules.moe_v2": ["initialize_moe_module"],
        "modules.custom_calls": ["CustomRMSNorm", "neuron_cumsum"],
        "modules.async_execution": ["causal_lm_async_execution"],
        "modules.autobucketing": ["generate_buckets"],
        "modules.padding": ["pad_tensor", "unpad_tensor"],
        "modules.checkpoint": ["load_state_dict", "save_state_dict_safetensors", "create_n_layer_checkpoint"],
        "modules.sliding_window.attention": ["flash_fwd"],
        "utils.hf_adapter": ["HuggingFaceGenerationAdapter", "load_pretrained_config"],
        "utils.accuracy": ["check_accuracy", "check_accuracy_logits", "generate_expected_logits"],
        "utils.benchmark": ["benchmark_sampling", "LatencyCollector"],
        "utils.snapshot": ["SnapshotOutputFormat", "register_snapshot_hooks"],
        "utils.tensor_capture_utils": ["capture_model_tensors", "get_available_modules"],
        "utils.tensor_replacement.registry": ["TensorReplacementRegistry"],
        "utils.kv_cache_reconstruct_utils": ["reconstruct_kv_cache"],
        "utils.debug_utils": ["capture_model_inputs"],
        "utils.runtime_env": ["set_env_vars"], "utils.compile_env": ["set_compile_env_vars"],
        "utils.distributed": ["get_init_world_size", "get_init_rank"], "utils.random": ["set_random_seed"],
        "utils.exceptions": ["LogitMatchingValidationError"], "utils.constants": ["MODEL_TYPES", "TEST_PROMPT"],
        "scripts.nxdi_distributed_launcher": ["main"], "inference_demo": ["main"],
        "models.llama.modeling_llama": ["NeuronLlamaForCausalLM", "NeuronLlamaModel", "LlamaInferenceConfig", "NeuronLlamaMLP", "NeuronLlamaAttention"],
        "models.diffusers.flux.application": ["NeuronFluxApplication"], "models.whisper.modeling_whisper": ["NeuronApplicationWhisper"],
        "experimental.functional": ["qkv_proj", "gated_mlp_fused", "tokengen_attention_megakernel_standard_kv"],
    }.items():
        m = importlib.import_module(P + mod)
        for n in names:
            assert hasattr(m, n), f"{mod}.{n}"


def test_diffusers_padder_activations_and_small_utils():
    """reference models/diffusers/{padder,activations}.py, utils/decorator_peeling.py, models/image_to_text_model_wrapper.py"""
    import functools
    import pytest
    from neuronx_distributed_inference_b200.models.diffusers.activations import FP32SiLU, NeuronGELU, get_activation
    from neuronx_distributed_inference_b200.models.diffusers.padder import MaybePadder, pad, pad_interleaved, pad_sizes, round_up_to_divisor
    from neuronx_distributed_inference_b200.models.image_to_text_model_wrapper import ImageToTextModelWrapper, VisionModelWrapper
    from neuronx_distributed_inference_b200.utils.decorator_peeling import peel_decorations
    assert round_up_to_divisor(24, 16) == 32 and round_up_to_divisor(32, 16) == 32
    assert pad_sizes((2, 3, 4), [0, 2], [5, 6]) == (0, 2, 0, 0, 0, 3) and pad_sizes((2, 3), 1, 5, left=True) == (2, 0, 0, 0)
    assert pad(torch.ones(2, 3), 1, 5).tolist() == [[1, 1, 1, 0, 0]] * 2 and pad(None, 0, 3) is None
    assert pad_interleaved(torch.tensor([1, 2, 3]), 0, 9, 1, 2).tolist() == [1, 0, 0, 2, 0, 0, 3, 0, 0]
    # 6 heads of width 2 padded to 8 heads for TP=2: every half gets 3 real heads + 1 zero head
    w = torch.arange(1, 13.0).view(12, 1).expand(12, 4).contiguous()
    p = MaybePadder(16, "interleaved", split_size=6, interleaved_factor=2)(w, 0)
    assert p.shape == (16, 4) and p[:, 0].tolist() == [1, 2, 3, 4, 5, 6, 0, 0, 7, 8, 9, 10, 11, 12, 0, 0]
    assert MaybePadder(5)(torch.ones(3, 2), 0).shape == (5, 2) and MaybePadder(5)(None, 0) is None
    g = NeuronGELU(8, 16, approximate="tanh")
    g.proj.weight.normal_()                       # parallel layers allocate with torch.empty: give the test defined values
    g.proj.bias.normal_()
    x = torch.randn(3, 8)
    assert torch.allclose(g(x), torch.nn.functional.gelu(torch.nn.functional.linear(x, g.proj.weight, g.proj.bias), approximate="tanh"))
    h = torch.randn(4, dtype=torch.bfloat16)
    assert FP32SiLU()(h).dtype == torch.bfloat16 and isinstance(get_activation("swish"), torch.nn.SiLU)
    with pytest.raises(ValueError):
        get_activation("nope")

    def deco(f):
        @functools.wraps(f)
        def inner(*a):
            return f(*a) + 1
        return inner

    def base(v):
        return v
    assert peel_decorations(deco(deco(base))) is base and deco(deco(base))(1) == 3
    assert ImageToTextModelWrapper.__name__ == "SubModelRunner" and VisionModelWrapper.__name__ == "EncoderRunner"


def test_learned_sink_input_processor_and_version(tmp_path):
    from neuronx_distributed_inference_b200 import _version
    from neuronx_distributed_inference_b200.modules.attention.sink import LearnedSink
    from neuronx_distributed_inference_b200.ops import reference as ref
    from neuronx_distributed_inference_b200.utils.input_processor import build_messages, prepare_generation_inputs_hf
    assert _version.__version__.count(".") == 2
    # the stand-alone sink module == the sink handling of the reference attention op
    sk = LearnedSink(1, 4)
    sk.sink.copy_(torch.tensor([0.5, -1.0, 2.0, 0.0]))
    q, k, v = torch.randn(2, 4, 3, 8), torch.randn(2, 4, 5, 8), torch.randn(2, 4, 5, 8)
    mask = torch.ones(2, 1, 3, 5, dtype=torch.bool)
    p = sk(q @ k.transpose(-1, -2) * 0.3)
    assert torch.allclose(p @ v, ref.attention_with_mask(q, k, v, mask, 0.3, sk.get_sink()), atol=1e-5) and (p.sum(-1) < 1).all()
    img = tmp_path / "x.jpg"
    img.write_bytes(b"\\xff\\xd8fake")
    msgs = build_messages("describe", [str(img), "https://host/y.png"])
    kinds = [c["type"] for c in msgs[0]["content"]]
    assert kinds == ["image", "image", "text"] and msgs[0]["content"][0]["url"].startswith("data:image/jpeg;base64,")

    class FakeProcessor:
        def apply_chat_template(self, messages, **kw):
            assert kw["add_generation_prompt"] and kw["return_dict"]
            return {"input_ids": torch.ones(1, 4, dtype=torch.long), "attention_mask": torch.ones(1, 4), "pixel_values": torch.zeros(1, 3, 2, 2), "unused": None}
    ids, m, extra = prepare_generation_inputs_hf("describe", str(img), FakeProcessor())
    assert ids.shape == (1, 4) and set(extra) == {"pixel_values"}


@pytest.mark.parametrize("past_len,seq_len", [(6, 1), (0, 5)])
def test_module_from_model_decoder_layer_orchestrator(past_len, seq_len, tmp_path):
    """reference module_test/{base_template,module_from_model_template}: ONE decoder layer of the HF model vs the same layer of the
    engine on shared random inputs and a shared pre-filled KV cache (decode step), or without a cache (prefill)."""
    import transformers as T
    from neuronx_distributed_inference_b200.config import load_pretrained_config
    from neuronx_distributed_inference_b200.models.llama.modeling_llama import NeuronLlamaForCausalLM as A
    from neuronx_distributed_inference_b200.module_test import DecoderLayerFromModelOrchestrator, OrchestratorConfig
    from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint
    ckpt = save_random_hf_checkpoint(T.LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                                   num_key_value_heads=2, vocab_size=128, max_position_embeddings=128), str(tmp_path / "m"), seed=1)
    hf = T.AutoModelForCausalLM.from_pretrained(ckpt, dtype=torch.float32)

    def factory(device):
        nc = A.get_neuron_config_cls()(batch_size=2, seq_len=32, max_context_length=16, torch_dtype="float32", on_cpu=device == "cpu")
        app = A(ckpt, A.get_config_cls()(nc, load_config=load_pretrained_config(ckpt)))
        return app.load(None, skip_warmup=True)
    orch = DecoderLayerFromModelOrchestrator(hf, factory, OrchestratorConfig(past_len=past_len, seq_len=seq_len, layer_idx=1, rtol=1e-4, atol=1e-4))
    rep = orch.run_validation(devices=("cpu",))
    assert rep["cpu"] < 1e-3


@pytest.mark.parametrize("D,Dp,split", [(100, 128, True), (80, 128, True), (48, 64, True), (96, 128, False)])
def test_zero_padded_heads_compute_the_same_attention(D, Dp, split):
    """modules/gqa.py head padding: Wqkv rows / Wo columns of every head zero-padded to a kernel width, rotary tables padded with the
    identity rotation, softmax scale of the ORIGINAL head size -> the attention block computes the same function."""
    import torch.nn.functional as F
    from neuronx_distributed_inference_b200.modules.gqa import _pad_heads
    from neuronx_distributed_inference_b200.ops import reference as ref
    torch.manual_seed(0)
    B, T, nq, nkv, H = 2, 5, 4, 2, 64
    wqkv, wo, x = torch.randn((nq + 2 * nkv) * D, H, dtype=torch.float64), torch.randn(H, nq * D, dtype=torch.float64), torch.randn(B, T, H, dtype=torch.float64)
    ang = torch.rand(B, T, D // 2, dtype=torch.float64) * 6.28

    def run(wqkv, wo, d, cos, sin):
        q, k, v = (x @ wqkv.t()).view(B, T, nq + 2 * nkv, d).split([nq, nkv, nkv], 2)
        q, k = ref.apply_rope(q, cos, sin, not split), ref.apply_rope(k, cos, sin, not split)
        return ref.attention_prefill(q, k, v, D ** -0.5, True).reshape(B, T, nq * d) @ wo.t()
    y0 = run(wqkv, wo, D, ang.cos(), ang.sin())
    y1 = run(_pad_heads(wqkv, D, Dp, 0, split), _pad_heads(wo, D, Dp, 1, split), Dp,
             F.pad(ang.cos(), (0, (Dp - D) // 2), value=1.0), F.pad(ang.sin(), (0, (Dp - D) // 2)))
    assert (y0 - y1).abs().max().item() < 1e-5 * y0.abs().max().item()      # (the reference ops evaluate softmax in fp32)


def test_module_test_template_tree_mlp_from_model(tmp_path):
    """reference module_test/{base_template,module_from_model_template}: the four-step adapters and the pairwise orchestrator — the MLP of
    decoder layer 1 cut out of a Hugging Face Llama and out of the engine application built from the same checkpoint."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from neuronx_distributed_inference_b200.config import NeuronConfig, load_pretrained_config
    from neuronx_distributed_inference_b200.models.llama.modeling_llama import LlamaInferenceConfig, NeuronLlamaForCausalLM
    from neuronx_distributed_inference_b200.module_test.base_template import ModuleAdapterBase, OrchestratorBase
    from neuronx_distributed_inference_b200.module_test.module_from_model_template import (
        MFMHFAdapter, MFMNxDICPUSingleRankAdapter, MFMOrchestratorBase, MFMOrchestratorConfig, build_prefixes_map,
        extract_subweights_by_prefixes)
    from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=160)
    ckpt = save_random_hf_checkpoint(cfg, str(tmp_path / "ck"), seed=1)
    assert build_prefixes_map(None, ["mlp"], 1, "model.layers") == {"mlp": "model.layers.1"}
    assert build_prefixes_map(["model", "model.layers"], ["rotary_emb", "self_attn"], 0, "x") == {"rotary_emb": "model", "self_attn": "model.layers.0"}
    sd = {"model.layers.1.mlp.up.weight": 1, "model.layers.1.mlp_other.w": 2, "model.layers.0.mlp.up.weight": 3}
    assert extract_subweights_by_prefixes({"mlp": "model.layers.1"}, sd) == {"mlp.up.weight": 1}

    def hf_forward(self, hidden_states):
        return self.mlp(hidden_states)

    def engine_forward(self, hidden_states):
        return self.mlp(hidden_states)

    def app_factory(path, device):
        nc = NeuronConfig(batch_size=2, seq_len=32, max_context_length=16, torch_dtype="float32", on_cpu=(device == "cpu"))
        app = NeuronLlamaForCausalLM(path, LlamaInferenceConfig(nc, load_config=load_pretrained_config(path)))
        return app.load(None, skip_warmup=True)
    conf = MFMOrchestratorConfig(batch_size=2, torch_dtype=torch.float32, hf_weight_ckpt_path=ckpt, layer_id=1, atol=1e-4, rtol=1e-4,
                                 prep_input_config=MFMOrchestratorConfig.PrepInputConfig(seq_len=3, hf_hidden_size=64))
    errs = MFMOrchestratorBase([MFMHFAdapter(hf_forward, LlamaForCausalLM, ["mlp"], layer_id=1),
                                MFMNxDICPUSingleRankAdapter(engine_forward, app_factory, ["mlp"], layer_id=1)], conf).run_validation()
    assert len(errs) == 1 and max(errs.values()) < 1e-4

    class Wrong(ModuleAdapterBase):          # a deliberately different module must be caught by the pairwise comparison
        def define_module_cls(self):
            class PassThrough(torch.nn.Module):
                def forward(self, hidden_states):
                    return hidden_states
            self.module_cls = PassThrough
    with pytest.raises(AssertionError, match="mismatch"):
        OrchestratorBase([MFMHFAdapter(hf_forward, LlamaForCausalLM, ["mlp"], layer_id=1), Wrong()], conf).run_validation()
    kv = OrchestratorBase([], MFMOrchestratorConfig(batch_size=2, torch_dtype=torch.float32, prep_input_config=conf.prep_input_config,
                                                    prep_kv_cache_config=MFMOrchestratorConfig.PrepKVCacheConfig(ctx_len=5, num_head=2, hf_head_hidden_size=16))
                          ).prepare_kv_cache_hf_format()
    assert kv[0].shape == (2, 2, 5, 16) and not torch.equal(kv[0], kv[1])


def test_reference_model_import_paths_and_helpers():
    """The reference's per-family module paths that this tree folds into one file each resolve, and their small helpers behave."""
    import importlib
    from types import SimpleNamespace
    base = "neuronx_distributed_inference_b200.models."
    for mod, names in {
        "qwen2_vl.modeling_qwen2_vl_text": ["NeuronQwen2VLTextModel", "NeuronQwen2VLTextForCausalLM"],
        "qwen2_vl.modeling_qwen2_vl_vision": ["NeuronQwen2VisionModel", "NeuronQwen2VLForImageEncoding", "PatchMerger"],
        "qwen3_vl.modeling_qwen3_vl_text": ["NeuronQwen3VLTextModel", "NeuronQwen3VLTextForCausalLM"],
        "qwen3_vl.modeling_qwen3_vl_vision": ["NeuronQwen3VLVisionModel", "NeuronQwen3VLForImageEncoding"],
        "pixtral.modeling_pixtral_vision": ["NeuronPixtralVisionModel", "NeuronPixtralForImageEncoding"],
        "mllama.modeling_mllama_vision": ["NeuronMllamaVisionModel"], "mllama.aspect_ratio_utils": ["get_all_supported_aspect_ratios"],
        "llama4.utils.input_processor": ["prepare_generation_inputs_hf"], "whisper.utils.state_dict": ["convert_hf_state_dict_to_neuron"],
        "gpt_oss.hf_configuration": ["GptOssConfig"], "deepseek.rope_util": ["DeepseekV3YarnRotaryEmbedding", "yarn_get_mscale"],
    }.items():
        m = importlib.import_module(base + mod)
        assert all(hasattr(m, n) for n in names), (mod, names)
    from neuronx_distributed_inference_b200.models.llama4.utils import encoder_utils as eu
    from neuronx_distributed_inference_b200.models.llama4.utils.layer_utils import is_after_nope_layer, is_before_nope_layer
    from neuronx_distributed_inference_b200.models.qwen2_vl.utils.vision_utils import calculate_max_grid_size, calculate_pixels_per_image
    from neuronx_distributed_inference_b200.models.qwen3_vl.utils.slicing import slice_by_image_hw
    from neuronx_distributed_inference_b200.models.whisper.utils.config import get_dims_from_config
    cfg = SimpleNamespace(no_rope_layers=[1, 1, 1, 0, 1])
    assert [is_before_nope_layer(cfg, i) for i in range(5)] == [False, False, True, False, False]
    assert [is_after_nope_layer(cfg, i) for i in range(5)] == [True, False, False, False, True]
    px, n = eu.pad_image_tensor(torch.ones(3, 2, 4, 4), 8)
    assert px.shape == (8, 2, 4, 4) and n == 3 and float(px[3:].abs().sum()) == 0 and eu.depad_output(px, n).shape[0] == 3
    mask = torch.tensor([[0, 1, 1, 0], [1, 0, 0, 0]], dtype=torch.bool)
    pos = eu.generate_positions_from_mask(mask)
    assert pos.tolist() == [1, 2, 4] and eu.pad_positions(pos, 5, 8).tolist() == [1, 2, 4, 8, 8]
    h = torch.zeros(2, 4, 3)
    out = eu.scatter_by_index_put(h, torch.arange(15.).view(5, 3), eu.pad_positions(pos, 5, 8))
    assert out[0, 1].tolist() == [0., 1., 2.] and out[1, 0].tolist() == [6., 7., 8.] and float(out[0, 0].abs().sum()) == 0
    assert eu.generate_llama4_vision_encoder_buckets(1, 16) == [1, 2, 4, 8, 16]
    assert calculate_pixels_per_image(640, 320) == (308 // 14) * (644 // 14) and calculate_max_grid_size(640, 320) == 46
    parts = slice_by_image_hw(torch.arange(24).view(24, 1), torch.tensor([[1, 2, 4], [1, 4, 4]]))
    assert [p.shape[0] for p in parts] == [8, 16]
    d = get_dims_from_config(SimpleNamespace(num_mel_bins=80, max_source_positions=1500, d_model=384, encoder_attention_heads=6, encoder_layers=4,
                                             vocab_size=51865, max_target_positions=448, decoder_attention_heads=6, decoder_layers=4))
    assert d.n_audio_state == 384 and d.n_text_layer == 4
