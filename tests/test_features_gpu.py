"""Serving features on the GPU (bf16, CUDA kernels + CUDA graphs): async decode, EAGLE / Medusa speculation, multimodal smoke."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
            vocab_size=512, head_dim=64)


def _mk(dtype="bfloat16", **kw):
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    return build_random_llama(TINY, batch_size=2, seq_len=128, max_context_length=32, device="cuda", dtype=dtype, seed=5, **kw)


def test_tree_speculation_exact_in_fp32_on_gpu():
    """Same device code path (indexing, KV compaction, masks) in fp32: must be exactly lossless; the bf16 tests below only
    differ by rounding between the multi-token and one-token kernels."""
    from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter
    torch.manual_seed(3)
    ids = torch.randint(1, 512, (2, 9))
    mask = torch.ones_like(ids)
    ref = HuggingFaceGenerationAdapter(_mk("float32")).generate(ids, attention_mask=mask, max_new_tokens=20)
    med = _mk("float32", is_medusa=True, num_medusa_heads=3, medusa_speculation_length=8, output_logits=True,
              medusa_tree=[[0], [1], [0, 0], [0, 1], [1, 0], [0, 0, 0]])
    seq = HuggingFaceGenerationAdapter(med).generate(ids, attention_mask=mask, max_new_tokens=20)
    assert torch.equal(seq[:, : ref.shape[1]], ref)
    tree = _mk("float32", speculation_length=4, enable_fused_speculation=True, enable_eagle_speculation=True,
               token_tree_config={"0": ["1", "2"], "1": ["3", "4"], "2": ["5"], "3": ["6"]},
               fused_draft=dict(hf=dict(num_hidden_layers=1), neuron=dict(is_eagle_draft=True)))
    seq = HuggingFaceGenerationAdapter(tree).generate(ids, attention_mask=mask, max_new_tokens=20)
    assert torch.equal(seq[:, : ref.shape[1]], ref)


def test_async_session_equals_sync_decode_gpu():
    from neuronx_distributed_inference_b200.modules.async_execution import causal_lm_async_execution
    app = _mk(async_mode=True)
    ids = torch.randint(1, 512, (2, 9))
    tok = app(ids, attention_mask=torch.ones_like(ids)).tokens.view(2, 1).cpu()
    pos = torch.full((2, 1), 9, dtype=torch.int32)
    got = causal_lm_async_execution(app, tok, pos, 12, depth=3)
    ref_app = _mk()
    t = ref_app(ids, attention_mask=torch.ones_like(ids)).tokens.view(2, 1).cpu()
    assert torch.equal(t, tok)
    ref = []
    for i in range(12):
        t = ref_app(t, position_ids=pos + i).tokens.view(2, 1).cpu()
        ref.append(t.view(-1))
    assert torch.equal(got, torch.stack(ref, 1))


@pytest.mark.parametrize("variant", ["eagle", "eagle_tree", "fused"])
def test_speculation_on_gpu_tracks_greedy(variant):
    """bf16 verification over k tokens runs different kernels (multi-token decode) than one-token decoding, so exact
    equality is not guaranteed at near-ties; require the run to work end to end and to agree on a long common prefix."""
    from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter
    ids = torch.randint(1, 512, (2, 9))
    mask = torch.ones_like(ids)
    ref = HuggingFaceGenerationAdapter(_mk()).generate(ids, attention_mask=mask, max_new_tokens=24)
    kw = dict(speculation_length=4, enable_fused_speculation=True)
    dn = {}
    if variant != "fused":
        kw["enable_eagle_speculation"] = True
        dn["is_eagle_draft"] = True
    if variant == "eagle_tree":
        kw["token_tree_config"] = {"0": ["1", "2"], "1": ["3", "4"], "2": ["5"], "3": ["6"]}
    app = _mk(fused_draft=dict(hf=dict(num_hidden_layers=1), neuron=dn), **kw)
    out = HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=24, return_dict_in_generate=True)
    seq = out.sequences
    n = min(seq.shape[1], ref.shape[1])
    same = (seq[:, :n] == ref[:, :n]).long().cumprod(-1).sum(-1)
    # the tree verify runs the masked (torch) attention path: more rounding differences vs the one-token kernels
    assert int(same.min()) >= 9 + (3 if variant == "eagle_tree" else 8) and int(same.max()) >= 9 + 12, (seq, ref)
    assert out.speculation_stats["steps"] > 0


def test_medusa_on_gpu():
    from neuronx_distributed_inference_b200.utils.hf_adapter import HuggingFaceGenerationAdapter
    ids = torch.randint(1, 512, (2, 9))
    mask = torch.ones_like(ids)
    ref = HuggingFaceGenerationAdapter(_mk()).generate(ids, attention_mask=mask, max_new_tokens=16)
    app = _mk(is_medusa=True, num_medusa_heads=3, medusa_speculation_length=8, output_logits=True,
              medusa_tree=[[0], [1], [0, 0], [0, 1], [1, 0], [0, 0, 0]])
    seq = HuggingFaceGenerationAdapter(app).generate(ids, attention_mask=mask, max_new_tokens=16)
    n = min(seq.shape[1], ref.shape[1])
    same = (seq[:, :n] == ref[:, :n]).long().cumprod(-1).sum(-1)
    assert int(same.min()) >= 9 + 3 and int(same.max()) >= 9 + 10


def test_flux_backbone_and_pipeline_gpu():
    from neuronx_distributed_inference_b200.config import NeuronConfig
    from neuronx_distributed_inference_b200.models.diffusers.flux.application import NeuronFluxApplication
    bb = dict(num_layers=2, num_single_layers=2, attention_head_dim=64, num_attention_heads=4, in_channels=64, joint_attention_dim=128,
              pooled_projection_dim=64, axes_dims_rope=(16, 24, 24), guidance_embeds=True)
    clip = dict(vocab_size=100, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, max_position_embeddings=16,
                hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2)
    t5 = dict(vocab_size=100, d_model=128, d_kv=32, d_ff=256, num_layers=2, num_heads=4, relative_attention_num_buckets=8,
              relative_attention_max_distance=16, layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu")
    vae = dict(latent_channels=16, out_channels=3, block_out_channels=(32, 64), layers_per_block=1, norm_num_groups=8,
               scaling_factor=0.36, shift_factor=0.11)
    app = NeuronFluxApplication(None, NeuronConfig(batch_size=1, torch_dtype="bfloat16"), bb, clip, t5, vae, height=64, width=64)
    app.load(random_weights=True)
    img = app(torch.randint(3, 99, (1, 8)), torch.randint(3, 99, (1, 6)), num_inference_steps=2)
    assert img.shape[0] == 1 and img.shape[1] == 3 and torch.isfinite(img).all()


def test_whisper_random_weights_gpu():
    from neuronx_distributed_inference_b200.config import NeuronConfig
    from neuronx_distributed_inference_b200.models.whisper.modeling_whisper import NeuronApplicationWhisper, WhisperInferenceConfig
    nc = NeuronConfig(batch_size=2, seq_len=64, max_context_length=16, torch_dtype="bfloat16")
    hf = dict(vocab_size=512, num_mel_bins=16, encoder_layers=2, encoder_attention_heads=4, decoder_layers=2, decoder_attention_heads=4,
              decoder_ffn_dim=512, encoder_ffn_dim=512, d_model=256, max_source_positions=64, max_target_positions=64, pad_token_id=0,
              eos_token_id=2, decoder_start_token_id=3, activation_function="gelu")
    cfg = WhisperInferenceConfig(nc, load_config=lambda c: [setattr(c, k, v) for k, v in hf.items()])
    app = NeuronApplicationWhisper("<random>", cfg)
    app.load(None, skip_warmup=True, random_weights=True)
    toks = app.generate(torch.randn(2, 16, 128), max_new_tokens=6, eos_token_id=-1)
    assert toks.shape == (2, 7)


def test_mixtral_decode_runs_under_cuda_graphs_with_moe_kernels():
    """MoE decode: routed experts through the moe_decode kernels, whole step replayed from a CUDA graph, tokens equal the
    eager (graph-free) run of the same model."""
    from neuronx_distributed_inference_b200 import ops
    from neuronx_distributed_inference_b200.models.mixtral.modeling_mixtral import NeuronMixtralForCausalLM
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    hf = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512,
              head_dim=64, num_local_experts=4, num_experts_per_tok=2)

    def mk(graphs):
        return build_random_llama(hf, batch_size=2, seq_len=128, max_context_length=32, device="cuda", dtype="bfloat16", seed=9,
                                  app_cls=NeuronMixtralForCausalLM, cuda_graphs=graphs)
    app = mk(True)
    assert app.model.graph_safe and app.token_generation_model.use_graphs
    ids = torch.randint(1, 512, (2, 9))

    def run(a):
        t = a(ids, attention_mask=torch.ones_like(ids)).tokens.view(2, 1).cpu()
        out = [t]
        for i in range(6):
            t = a(t, position_ids=torch.full((2, 1), 9 + i, dtype=torch.int32)).tokens.view(2, 1).cpu()
            out.append(t)
        return torch.cat(out, 1)
    before = ops.stats["moe_decode"]
    got = run(app)
    assert ops.stats["moe_decode"] > before and len(app.token_generation_model._graphs) > 0
    assert torch.equal(got, run(mk(False)))


def test_gpt_oss_and_dbrx_run_under_cuda_graphs_with_grouped_moe():
    """Families whose experts the moe_decode kernels do not cover (GPT-OSS: biases + clamped SwiGLU) or whose prefill used to be eager
    (every MoE): routed experts through the grouped tcgen05 GEMMs at every token count -> prefill AND decode replay from CUDA graphs,
    tokens equal the graph-free run."""
    from neuronx_distributed_inference_b200 import ops
    from neuronx_distributed_inference_b200.models.dbrx.modeling_dbrx import NeuronDbrxForCausalLM
    from neuronx_distributed_inference_b200.models.gpt_oss.modeling_gpt_oss import NeuronGptOssForCausalLM
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    cases = [
        (NeuronGptOssForCausalLM, dict(hidden_size=256, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                       vocab_size=512, head_dim=64, num_local_experts=4, num_experts_per_tok=2, sliding_window=16,
                                       layer_types=["sliding_attention", "full_attention"])),
        (NeuronDbrxForCausalLM, dict(d_model=256, n_heads=4, n_layers=2, max_seq_len=256, vocab_size=512,
                                     attn_config=dict(kv_n_heads=2, clip_qkv=8.0, rope_theta=10000.0),
                                     ffn_config=dict(ffn_hidden_size=256, moe_num_experts=4, moe_top_k=2, hidden_size=256))),
    ]
    for cls, hf in cases:
        def mk(graphs):
            return build_random_llama(hf, batch_size=2, seq_len=128, max_context_length=32, device="cuda", dtype="bfloat16", seed=9,
                                      app_cls=cls, cuda_graphs=graphs)
        app = mk(True)
        assert app.model.graph_safe and app.token_generation_model.use_graphs, cls.__name__
        ids = torch.randint(1, 512, (2, 20))

        def run(a):
            t = a(ids, attention_mask=torch.ones_like(ids)).tokens.view(2, 1).cpu()
            out = [t]
            for i in range(6):
                t = a(t, position_ids=torch.full((2, 1), 20 + i, dtype=torch.int32)).tokens.view(2, 1).cpu()
                out.append(t)
            return torch.cat(out, 1)
        before = ops.stats["moe_grouped"]
        got = run(app)
        assert ops.stats["moe_grouped"] > before and len(app.token_generation_model._graphs) > 0, cls.__name__
        assert torch.equal(got, run(mk(False))), cls.__name__


@pytest.mark.parametrize("cfg", [
    dict(hidden_size=1024, intermediate_size=2816, num_hidden_layers=3, num_attention_heads=8, num_key_value_heads=2, head_dim=128),
    dict(hidden_size=2048, intermediate_size=1792, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=1, head_dim=128),
    dict(hidden_size=1024, intermediate_size=2048, num_hidden_layers=2, num_attention_heads=16, num_key_value_heads=16, head_dim=64),
])
def test_persistent_decode_step_matches_layer_by_layer(cfg):
    """csrc/decode_step.cu (all layers of a decode step in one persistent launch) vs the per-kernel path: same logits."""
    from neuronx_distributed_inference_b200 import ops
    from neuronx_distributed_inference_b200.runtime import decode_step
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    torch.manual_seed(0)
    app = build_random_llama(dict(cfg, vocab_size=4096), batch_size=2, seq_len=256, max_context_length=128, device="cuda",
                             output_logits=True)
    ids = torch.randint(0, 4096, (2, 100))

    def run(enabled):
        decode_step._ENABLED = enabled
        app.reset()
        out = app(ids, attention_mask=torch.ones_like(ids))
        tok = out.tokens
        pos = torch.full((2, 1), 100, dtype=torch.int32)
        logits = []
        for _ in range(6):
            out = app(tok.view(2, 1).cpu(), position_ids=pos)
            logits.append(out.logits.float().clone())
            tok = out.tokens
            pos = pos + 1
        torch.cuda.synchronize()
        return torch.stack(logits)
    try:
        n0 = ops.stats["decode_step"]
        a = run(True)
        assert ops.stats["decode_step"] > n0, "persistent decode-step kernel was not used"
        b = run(False)
    finally:
        decode_step._ENABLED = False
    rel = ((a - b).norm() / b.norm()).item()
    assert rel < 2e-2, rel
    assert (a.argmax(-1) == b.argmax(-1)).float().mean().item() > 0.9


@pytest.mark.gpu
@pytest.mark.parametrize("head_dim,heads,kv", [(100, 4, 4), (80, 4, 2), (48, 8, 2)])
def test_odd_head_dims_run_on_padded_kernels(head_dim, heads, kv, monkeypatch):
    """Head sizes the attention kernels are not specialised for (open_llama_3b: 100, ...) are stored zero-padded to 64 / 128 channels on
    the CUDA bf16 path (modules/gqa.py, attention_base.py).  Same weights, padding off (PyTorch composite attention) vs on (kernels):
    same logits up to bf16 rounding, and the padded model is graph-safe."""
    from neuronx_distributed_inference_b200 import ops
    from neuronx_distributed_inference_b200.utils.testing import build_random_llama
    hf = dict(hidden_size=heads * head_dim, intermediate_size=512, num_hidden_layers=2, num_attention_heads=heads,
              num_key_value_heads=kv, head_dim=head_dim, vocab_size=512)
    ids = torch.randint(1, 512, (2, 12))
    mask = torch.ones_like(ids)

    feed = [torch.randint(1, 512, (2, 1)) for _ in range(4)]      # teacher-forced decode inputs (greedy feedback would amplify near-ties)

    def logits_of(app):
        app.reset()
        out = app(ids, attention_mask=mask)
        lg = [out.logits[:, -1].float().clone()]
        pos = torch.full((2, 1), 12, dtype=torch.int32)
        for i, tok in enumerate(feed):
            lg.append(app(tok, position_ids=pos + i).logits[:, -1].float().clone())
        return torch.stack(lg)

    def build(pad):
        monkeypatch.setenv("NXDI_B200_PAD_HEAD_DIM", "1" if pad else "0")
        app = build_random_llama(hf, batch_size=2, seq_len=64, max_context_length=16, device="cuda", dtype="bfloat16", seed=11,
                                 output_logits=True)
        attn = app.model.layers[0].self_attn
        assert attn.head_dim == ((64 if head_dim < 64 else 128) if pad else head_dim) and attn.logical_head_dim == head_dim
        return app
    app0 = build(False)
    l0 = logits_of(app0)
    sd = {k: v.clone() for k, v in app0.model.state_dict().items()}
    app1 = build(True)
    assert app1.model.graph_safe
    # copy the unpadded weights into the padded layout through the layers' own shard functions
    with torch.no_grad():
        for name, p_ in app1.model.named_parameters():
            src = sd[name]
            p_.copy_(p_.shard_fn(src.cpu(), 0).to(p_.device) if (src.shape != p_.shape and hasattr(p_, "shard_fn")) else src)
    n0 = ops.stats["rope_attn_decode"]
    l1 = logits_of(app1)
    assert ops.stats["rope_attn_decode"] > n0
    rel = ((l1 - l0).norm() / l0.norm()).item()
    assert rel < 3e-2, rel
