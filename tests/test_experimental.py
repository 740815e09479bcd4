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


def test_functional_llama4_matches_application(tmp_path):
    """Functional Llama-4 (iRoPE: chunked-attention RoPE layers + NoPE global layers, sigmoid top-1 MoE with shared expert) generates
    the same tokens as the Hugging-Face-validated application on the same weights (reference experimental/models/llama4/model.py)."""
    import transformers as T
    from neuronx_distributed_inference_b200.config import load_pretrained_config
    from neuronx_distributed_inference_b200.experimental.core import generate
    from neuronx_distributed_inference_b200.experimental.models.config import Config
    from neuronx_distributed_inference_b200.experimental.models.llama4.model import Llama4
    from neuronx_distributed_inference_b200.models.llama4.modeling_llama4_text import NeuronLlama4TextForCausalLM as A
    from neuronx_distributed_inference_b200.utils.testing import save_random_hf_checkpoint
    hf_cfg = T.Llama4TextConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2,
                                vocab_size=160, max_position_embeddings=256, head_dim=16, num_local_experts=4, num_experts_per_tok=1,
                                intermediate_size_mlp=128, interleave_moe_layer_step=2, attention_chunk_size=8, no_rope_layers=[1, 1, 1, 0],
                                use_qk_norm=True, attn_temperature_tuning=True, floor_scale=4, attn_scale=0.1)
    ckpt = save_random_hf_checkpoint(hf_cfg, str(tmp_path / "l4"), seed=2)
    nc = A.get_neuron_config_cls()(batch_size=2, seq_len=48, max_context_length=24, torch_dtype="float32", on_cpu=True)
    app = A(ckpt, A.get_config_cls()(nc, load_config=load_pretrained_config(ckpt)))
    app.load(None, skip_warmup=True)
    ids = torch.randint(1, 160, (2, 13))
    mask = torch.ones_like(ids)
    mask[1, 10:] = 0
    ref = HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=8)
    cfg = Config.from_hf(hf_cfg.to_dict(), max_batch_size=2, max_seq_len=48, dtype=torch.float32)
    assert cfg.nope_layers == [3] and cfg.moe_layers == [1, 3] and cfg.ffn_dim == 128 and cfg.moe_ffn_dim == 128
    model = Llama4(cfg, dict(app.model.state_dict()), torch.device("cpu"))
    got = generate(model, ids, mask, max_new_tokens=8)
    for b in range(2):
        n = int(mask[b].sum()) + 8
        assert got[b, :n].tolist() == ref[b, :n].tolist()


def test_llama3_tokenizer_and_chat_format(tmp_path):
    """tiktoken BPE over a rank file + the Llama-3 special tokens and dialog framing (reference experimental/models/llama3/tokenizer.py)."""
    import base64
    from neuronx_distributed_inference_b200.experimental.models.llama3.tokenizer import ChatFormat, Tokenizer
    ranks = [bytes([i]) for i in range(256)] + [b"he", b"ll", b"hell", b"hello", b" w", b"or", b" wor", b"ld", b" world"]
    path = tmp_path / "tokenizer.model"
    path.write_text("\n".join(f"{base64.b64encode(tok).decode()} {i}" for i, tok in enumerate(ranks)))
    tk = Tokenizer(str(path))
    assert tk.n_words == len(ranks) + 256 and tk.bos_id == len(ranks) and tk.special_tokens["<|eot_id|>"] == len(ranks) + 9
    ids = tk.encode("hello world", bos=True, eos=True)
    assert ids == [tk.bos_id, 259, 264, tk.eos_id] and tk.decode(ids[1:-1]) == "hello world"
    long = "a" * 60_000 + " " + "b" * 10
    assert tk.decode(tk.encode(long, bos=False, eos=False)) == long            # sliced encoding of a long no-whitespace run
    chat = ChatFormat(tk)
    p = chat.encode_dialog_prompt([{"role": "system", "content": "be brief"}, {"role": "user", "content": " hello world "}])
    sh, eh, eot = (tk.special_tokens[t] for t in ("<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"))
    assert p[0] == tk.bos_id and p.count(sh) == 3 and p.count(eh) == 3 and p.count(eot) == 2 and p[-1] != eot
    assert tk.decode(p).endswith("<|start_header_id|>assistant<|end_header_id|>\n\n") and "hello world<|eot_id|>" in tk.decode(p)


def test_experimental_core_build_flow_processor_and_generate(tmp_path):
    """The reference's experimental flow end to end on the functional Llama-3: YAML with per-tag overrides -> bucketed build ->
    BucketingProcessor -> list-style greedy generate (experimental/core/**)."""
    from neuronx_distributed_inference_b200.experimental.core.build_flow import build_for_bucketing_on_seq_len
    from neuronx_distributed_inference_b200.experimental.core.config import (get_config_for_model_tag, load_neuron_config,
                                                                             parse_config_with_model_tags_overrides)
    from neuronx_distributed_inference_b200.experimental.core.functions import build
    from neuronx_distributed_inference_b200.experimental.core.generate import GenerateResult, generate
    from neuronx_distributed_inference_b200.experimental.core.pad import pad_at_end, pad_to_shape
    from neuronx_distributed_inference_b200.experimental.core.processor import (BucketingProcessor, collect_buckets,
                                                                                select_smallest_bucket)
    from neuronx_distributed_inference_b200.experimental.models.llama3.model import Llama3, Llama3Args
    # ---- YAML: defaults + per-tag overrides (dict sections merge, compiler_args extend, other lists replace)
    y = tmp_path / "cfg.yaml"
    y.write_text("""
model: {name: llama3}
build: {batch_size: 2, compiler_args: ["--a"], buckets: [16, 32]}
attention: {try_using_kernel: true}
dtype: ${torch_dtype:bfloat16}
config_override:
  - model_tags: [prefill_16, prefill_32]
    build: {compiler_args: ["--b"], buckets: [8]}
  - model_tags: [decode_32]
    attention: {try_using_kernel: false}
    extra: {x: 1}
""")
    cfg = load_neuron_config(str(y))
    assert cfg.default.dtype is torch.bfloat16 and cfg.model.name == "llama3" and "config_override" not in cfg.default
    p16 = get_config_for_model_tag(cfg, "prefill_16")
    assert p16.build.compiler_args == ["--a", "--b"] and p16.build.buckets == [8] and p16.build.batch_size == 2
    assert p16.attention.try_using_kernel is True and cfg.default.build.compiler_args == ["--a"]
    d32 = get_config_for_model_tag(cfg, "decode_32")
    assert d32.attention.try_using_kernel is False and d32.extra.x == 1 and d32.build.compiler_args == ["--a"]
    assert get_config_for_model_tag(cfg, "unknown_tag") is cfg.default
    assert parse_config_with_model_tags_overrides(cfg.__class__({"model": {}, "config_override": []})).default == {}
    # ---- padding helpers
    t = torch.arange(6).view(2, 3)
    assert pad_at_end(t, 1, 5, value=9).tolist() == [[0, 1, 2, 9, 9], [3, 4, 5, 9, 9]] and pad_to_shape(t, (3, 4), value=0).shape == (3, 4)
    assert pad_to_shape(t, (2, 3)) is t
    # ---- build + processor + generate on the functional model, against the application's tokens
    tiny = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                vocab_size=128, head_dim=16, rope_theta=10000.0)
    app = build_random_llama(tiny, batch_size=2, seq_len=48, max_context_length=16, device="cpu", dtype="float32", seed=4)
    args = Llama3Args(dim=64, n_layers=2, n_heads=4, n_kv_heads=2, vocab_size=128, ffn_dim=128, norm_eps=1e-5, rope_theta=10000.0,
                      max_batch_size=2, max_seq_len=48, dtype=torch.float32)
    model = Llama3(args, dict(app.model.state_dict()), torch.device("cpu"))
    built = build_for_bucketing_on_seq_len(model, world_size=1, batch_size=2, prefill_buckets=[8, 16], decode_buckets=[16, 32, 48])
    assert sorted(built.reserved_example_inputs) == ["decode_16", "decode_32", "decode_48", "prefill_16", "prefill_8"]
    buckets, table = collect_buckets(built.reserved_example_inputs)
    assert table == {"prefill": [8, 16], "decode": [16, 32, 48]} and select_smallest_bucket(buckets["decode"], table["decode"], 17)[2].shape == (2, 32)
    proc = BucketingProcessor(built, pad_token_id=0)
    prompts = [[5, 9, 33, 7, 21, 3, 90], [11, 2, 64, 8, 19]]
    res = generate(proc, 7 + 6 - 1, [list(p) for p in prompts], stop_tokens=[], pad_token=0)
    assert isinstance(res, GenerateResult) and res.logits is None
    ids = torch.tensor([prompts[0], prompts[1] + [0, 0]])
    mask = torch.tensor([[1] * 7, [1] * 5 + [0, 0]])
    ref = HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=6)
    for b in range(2):
        n = len(prompts[b])
        assert res.prompt_tokens[b][: n + 5] == ref[b, : n + 5].tolist()
    single = build(model, batch_size=2, sequence_length=16)
    assert sorted(single.reserved_example_inputs) == ["decode", "prefill"] and single.reserved_example_inputs["decode"][0].shape == (2, 1)


def test_experimental_functional_module_tree_and_meshes():
    """The reference's functional import paths resolve, and the context- / data-parallel mesh helpers compute the documented layouts."""
    from neuronx_distributed_inference_b200.experimental.functional.attention.causal_attention_functions import qkv_proj  # noqa: F401
    from neuronx_distributed_inference_b200.experimental.functional.attention.data_parallel import split_input_for_data_parallel
    from neuronx_distributed_inference_b200.experimental.functional.attention.output_projection import o_proj_kernel_unreduced  # noqa: F401
    from neuronx_distributed_inference_b200.experimental.functional.attention.tokengen_attention.tokengen_attention_block_kv import (  # noqa: F401
        tokengen_attention_megakernel_block_kv)
    from neuronx_distributed_inference_b200.experimental.functional.attention.tokengen_attention.tokengen_attention_standard_kv import (  # noqa: F401
        tokengen_attention_megakernel_standard_kv)
    from neuronx_distributed_inference_b200.experimental.functional.ffn.mlp import gated_mlp_kernel_unreduced  # noqa: F401
    from neuronx_distributed_inference_b200.experimental.functional.moe.tokengen_moe.tokengen_moe_forward_all_experts import (  # noqa: F401
        tokengen_moe_megakernel_forward_all_experts)
    from neuronx_distributed_inference_b200.experimental.functional.norm.norm_functions import rmsnorm
    from neuronx_distributed_inference_b200.experimental.functional.parallel.tensor_ops import split_along_dim
    from neuronx_distributed_inference_b200.experimental.functional.pg import (get_context_parallel_cp_mesh, get_context_parallel_tp_mesh,
                                                                               get_cp_rank, get_dp_rank)
    from neuronx_distributed_inference_b200.experimental.functional.qkv.qkv import qkv_kernel  # noqa: F401
    assert get_context_parallel_tp_mesh(8, 2) == [[0, 1, 2, 3], [4, 5, 6, 7]]
    assert get_context_parallel_cp_mesh(8, 2) == [[0, 4], [1, 5], [2, 6], [3, 7]]
    assert get_cp_rank(torch.tensor([0, 3, 4, 7]), 8, 2).tolist() == [0, 0, 1, 1] and int(get_dp_rank(torch.tensor(5), 8, 4)) == 2
    x = torch.arange(24).view(4, 6)
    assert split_along_dim(x, 1, 2, 3).tolist() == [[4, 5], [10, 11], [16, 17], [22, 23]]
    assert split_input_for_data_parallel(x, 0, 8, 2, torch.tensor(6)).tolist() == x[2:].tolist()
    w = torch.rand(6) + 0.5
    xf = torch.randn(2, 6)
    assert torch.allclose(rmsnorm(xf, w, 1e-6), xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w, atol=1e-5)
