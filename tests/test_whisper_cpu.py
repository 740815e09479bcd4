"""Whisper: encoder output, teacher-forced decoder logits (prefill + decode with cached cross K/V) and greedy generation vs HF."""
import torch

from neuronx_distributed_inference_b200.config import load_pretrained_config
from neuronx_distributed_inference_b200.utils.constants import get_model_cls


def test_whisper_matches_hf(tmp_path):
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    torch.manual_seed(0)
    cfg = WhisperConfig(vocab_size=120, num_mel_bins=8, encoder_layers=2, encoder_attention_heads=2, decoder_layers=2,
                        decoder_attention_heads=2, decoder_ffn_dim=64, encoder_ffn_dim=64, d_model=32, max_source_positions=20,
                        max_target_positions=40, pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=3,
                        suppress_tokens=None, begin_suppress_tokens=None)
    hf = WhisperForConditionalGeneration(cfg).eval()
    ckpt = str(tmp_path / "whisper")
    hf.save_pretrained(ckpt)
    cls = get_model_cls("whisper", "speech-to-text")
    nc = cls.get_neuron_config_cls()(batch_size=2, seq_len=32, max_context_length=16, torch_dtype="float32", on_cpu=True, output_logits=True)
    app = cls(ckpt, cls.get_config_cls()(nc, load_config=load_pretrained_config(ckpt)))
    app.load(None, skip_warmup=True)
    mel = torch.randn(2, 8, 40)
    dec = torch.randint(3, 120, (2, 5))
    with torch.no_grad():
        exp = hf(input_features=mel, decoder_input_ids=dec)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()   # noqa: E731
    assert rel(app.encode(mel), exp.encoder_last_hidden_state) < 1e-4
    out = app(dec, input_features=mel)
    assert rel(out.logits[:, -1], exp.logits[:, -1]) < 2e-4
    nxt = exp.logits[:, -1].argmax(-1)
    with torch.no_grad():
        exp2 = hf(input_features=mel, decoder_input_ids=torch.cat([dec, nxt.view(2, 1)], 1)).logits[:, -1]
    out2 = app(nxt.view(2, 1), position_ids=torch.full((2, 1), 5, dtype=torch.int32))
    assert rel(out2.logits[:, -1], exp2) < 2e-4
    # greedy generation
    start = torch.full((2, 1), 3)
    ref = start.clone()
    with torch.no_grad():        # plain greedy loop (HF's Whisper generate adds task-specific logits processors)
        for _ in range(8):
            ref = torch.cat([ref, hf(input_features=mel, decoder_input_ids=ref).logits[:, -1].argmax(-1, keepdim=True)], 1)
    got = app.generate(mel, start, max_new_tokens=8, eos_token_id=-1)
    assert torch.equal(got, ref)
