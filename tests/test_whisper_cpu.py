"""Whisper: encoder output, teacher-forced decoder logits (prefill + decode with cached cross K/V) and greedy generation vs HF."""
import torch

from neuronx_distributed_inference_b200.utils.testing import perturb_constant_vectors  # noqa: E402

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
    perturb_constant_vectors(hf)
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


def test_whisper_decoding_rules_and_fallback(tmp_path):
    """Timestamp grammar vs the Hugging Face logits processor on random histories; suppress filters; decode() == generate() for plain
    greedy decoding; the fallback loop moves to the next temperature when the thresholds reject a result."""
    from types import SimpleNamespace
    from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor
    from neuronx_distributed_inference_b200.models.whisper.utils.decoding import (DecodingOptions, apply_timestamp_rules, compression_ratio,
                                                                                 decode, suppress_tokens)
    V, notime, eos = 120, 99, 90
    hf_proc = WhisperTimeStampLogitsProcessor(SimpleNamespace(no_timestamps_token_id=notime, eos_token_id=eos, bos_token_id=eos,
                                                              max_initial_timestamp_index=5, _detect_timestamp_from_logprob=True), begin_index=2)
    g = torch.Generator().manual_seed(0)
    for n in (0, 1, 2, 3, 6):
        for _ in range(6):
            hist = torch.randint(0, V, (3, n), generator=g)
            hist[hist == notime] = 5
            for b in range(3):                       # valid histories only: sort the timestamp tokens so that they never decrease
                ts = hist[b] >= notime + 1
                hist[b, ts] = hist[b, ts].sort().values
            scores = torch.randn(3, V, generator=g) * 3
            exp = hf_proc(torch.cat([torch.zeros(3, 2, dtype=torch.long), hist], 1), scores)
            got = apply_timestamp_rules(scores, hist, notime, eos, 5)
            assert torch.equal(torch.isinf(got), torch.isinf(exp)) and torch.allclose(got[~torch.isinf(got)], exp[~torch.isinf(exp)])
    s = suppress_tokens(torch.zeros(2, 10), [1, 7])
    assert torch.isinf(s[:, [1, 7]]).all() and torch.isfinite(s[:, 0]).all()
    assert compression_ratio([7] * 200) > 5 and compression_ratio(list(range(50))) < 2.4
    # end to end on a tiny random Whisper
    import transformers as T
    from neuronx_distributed_inference_b200.config import NeuronConfig, load_pretrained_config
    from neuronx_distributed_inference_b200.models.whisper.modeling_whisper import NeuronApplicationWhisper as A
    cfg = T.WhisperConfig(vocab_size=120, d_model=32, encoder_layers=1, decoder_layers=1, encoder_attention_heads=2, decoder_attention_heads=2,
                          encoder_ffn_dim=64, decoder_ffn_dim=64, num_mel_bins=8, max_source_positions=20, max_target_positions=32,
                          decoder_start_token_id=1, eos_token_id=90, pad_token_id=90, bos_token_id=90, suppress_tokens=[], begin_suppress_tokens=[])
    torch.manual_seed(0)
    hf = T.WhisperForConditionalGeneration(cfg).eval()
    ckpt = str(tmp_path / "w")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    nc = NeuronConfig(batch_size=2, seq_len=32, max_context_length=8, torch_dtype="float32", on_cpu=True, output_logits=True)
    app = A(ckpt, A.get_config_cls()(nc, load_config=load_pretrained_config(ckpt)))
    app.load(None, skip_warmup=True)
    feats = torch.randn(2, 8, 40)
    prompt = torch.full((2, 2), 1, dtype=torch.long)
    plain = app.generate(feats, decoder_input_ids=prompt, max_new_tokens=10, eos_token_id=90)
    res = decode(app, feats, prompt, DecodingOptions(max_new_tokens=10, temperatures=(0.0,), eos_token_id=90, logprob_threshold=None,
                                                     compression_ratio_threshold=None))
    for b in range(2):
        exp = plain[b, 2:].tolist()
        exp = exp[: exp.index(90)] if 90 in exp else exp
        assert res[b].tokens == exp and res[b].temperature == 0.0 and res[b].avg_logprob <= 0
    # an impossible log-prob threshold rejects every temperature: the result of the LAST temperature is returned
    res = decode(app, feats, prompt, DecodingOptions(max_new_tokens=6, temperatures=(0.0, 0.7), eos_token_id=90, logprob_threshold=0.0))
    assert all(r.temperature == 0.7 for r in res)
    # timestamps on: the first generated token is a timestamp within the allowed initial range
    res = decode(app, feats, prompt, DecodingOptions(max_new_tokens=6, temperatures=(0.0,), eos_token_id=90, without_timestamps=False,
                                                     no_timestamps_token_id=99, max_initial_timestamp_index=3, logprob_threshold=None))
    assert all(100 <= r.tokens[0] <= 103 for r in res if r.tokens)
