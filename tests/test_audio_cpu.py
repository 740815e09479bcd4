"""Audio encoders on the encoder-application path: wav2vec 2.0 frame classification (LaughterSegmentation) vs Hugging Face."""
import pytest
import torch

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
