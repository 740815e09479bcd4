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
