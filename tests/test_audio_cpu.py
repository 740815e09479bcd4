"""Audio encoders on the encoder-application path: wav2vec 2.0 frame classification (LaughterSegmentation) vs Hugging Face."""
import pytest
import torch

from neuronx_distributed_inference_b200.utils.testing import perturb_constant_vectors  # noqa: E402

from neuronx_distributed_inference_b200.config import NeuronConfig, load_pretrained_config


@pytest.mark.parametrize("stable", [True, False])
def test_wav2vec2_frame_classifier_matches_hf(stable, tmp_path):
    import transformers as T
    from neuronx_distributed_inference_b200.contrib.models.wav2vec2 import NeuronWav2Vec2ForAudioFrameClassification as A
    torch.manual_seed(0)
    cfg = T.Wav2Vec2Config(hidden_size=32, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64, conv_dim=(16, 16, 16),
                           conv_stride=(5, 2, 2), conv_kernel=(10, 3, 3), num_conv_pos_embeddings=8, num_conv_pos_embedding_groups=4,
                           do_stable_layer_norm=stable, feat_extract_norm="layer" if stable else "group", conv_bias=stable, num_labels=2,
                           vocab_size=32, use_weighted_layer_sum=not stable)
    hf = T.Wav2Vec2ForAudioFrameClassification(cfg).eval()
    ckpt = str(tmp_path / "w2v")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    nc = NeuronConfig(batch_size=4, torch_dtype="float32", on_cpu=True, buckets=[4])
    app = A(ckpt, A.get_config_cls()(nc, load_config=load_pretrained_config(ckpt)))
    app.load(None, skip_warmup=True)
    wav = torch.randn(3, 2000)                                            # 3 clips, padded to the batch bucket of 4 inside the runner
    with torch.no_grad():
        exp = hf(wav).logits
    got = app(wav)[:3]
    assert got.shape == exp.shape == (3, app.model.num_frames(2000), 2)
    assert ((got - exp).norm() / exp.norm()) < 1e-4
    assert app.encoder_model.tag == "audio_encoder_model"


def test_qwen2_5_omni_thinker_audio_text_matches_hf(tmp_path):
    """Audio tower (windowed Whisper-style encoder) + Qwen2 decoder: two audios of different length (one spanning two windows) scattered
    into the prompts; prefill and a decode step against the Hugging Face thinker."""
    import json
    import os
    import transformers as T
    from safetensors.torch import save_file
    from transformers.models.qwen2_5_omni import modeling_qwen2_5_omni as M
    from neuronx_distributed_inference_b200.contrib.models.qwen2_5_omni import NeuronQwen2_5OmniAudioEncoder as Enc
    from neuronx_distributed_inference_b200.contrib.models.qwen2_5_omni import NeuronQwen2_5OmniThinkerForCausalLM as A
    torch.manual_seed(0)
    tc = T.Qwen2_5OmniThinkerConfig(
        text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=200,
                         max_position_embeddings=256, rope_parameters=dict(rope_type="default", mrope_section=[2, 3, 3], rope_theta=10000.0)),
        audio_config=dict(d_model=32, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=64, num_mel_bins=8, output_dim=64, n_window=8,
                          max_source_positions=64),
        vision_config=dict(depth=1, hidden_size=16, intermediate_size=32, num_heads=2, out_hidden_size=64, patch_size=4, spatial_merge_size=2,
                           temporal_patch_size=2),
        audio_token_index=150, image_token_index=151, video_token_index=152, audio_start_token_id=153, audio_end_token_id=154,
        position_id_per_seconds=25, seconds_per_chunk=2, vision_start_token_id=155, vision_end_token_id=156, vision_token_id=157,
        audio_token_id=150, image_token_id=151, video_token_id=152)
    hf = M.Qwen2_5OmniThinkerForConditionalGeneration(tc).eval()
    dst = str(tmp_path / "omni")
    os.makedirs(dst)
    save_file({"thinker." + k: v.clone().contiguous() for k, v in hf.state_dict().items()}, os.path.join(dst, "model.safetensors"))
    json.dump({"model_type": "qwen2_5_omni_test", "thinker_config": tc.to_dict()}, open(os.path.join(dst, "config.json"), "w"))
    nc = NeuronConfig(batch_size=2, seq_len=64, max_context_length=32, torch_dtype="float32", on_cpu=True, output_logits=True)
    app = A(dst, A.get_config_cls()(nc, load_config=load_pretrained_config(dst)))
    app.load(None, skip_warmup=True)
    feats = torch.randn(2, 8, 30)
    fmask = torch.zeros(2, 30, dtype=torch.long)
    fmask[0, :30] = 1                                         # 30 frames: windows of 16 + 14  -> 15 tokens -> 7 after pooling
    fmask[1, :11] = 1                                         # 11 frames: one window          -> 6 tokens  -> 3 after pooling
    n_out = Enc.output_lengths(fmask.sum(1))[1].tolist()
    assert n_out == [7, 3]
    ids = torch.randint(1, 140, (2, 16))
    ids[0, 2:9] = 150
    ids[1, 4:7] = 150
    mask = torch.ones_like(ids)
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, input_features=feats, feature_attention_mask=fmask)
        aud = hf.get_audio_features(feats, feature_attention_mask=fmask).last_hidden_state
    got_aud = app.encode_images(feats, feature_attention_mask=fmask)
    assert got_aud.shape == aud.shape == (10, 64) and ((got_aud - aud).norm() / aud.norm()) < 1e-4
    out = app(ids, attention_mask=mask, input_features=feats, feature_attention_mask=fmask)
    assert ((out.logits[:, -1] - exp.logits[:, -1]).norm() / exp.logits[:, -1].norm()) < 2e-4
    nxt = exp.logits[:, -1].argmax(-1)
    ids2 = torch.cat([ids, nxt.view(2, 1)], 1)
    with torch.no_grad():
        exp2 = hf(input_ids=ids2, attention_mask=torch.ones_like(ids2), input_features=feats, feature_attention_mask=fmask).logits[:, -1]
    out2 = app(nxt.view(2, 1), position_ids=torch.full((2, 1), 16, dtype=torch.int32))
    assert ((out2.logits[:, -1] - exp2).norm() / exp2.norm()) < 2e-4
