"""Image-to-text models on CPU vs Hugging Face (tiny random weights): vision tower output, prefill logits with images
scattered into the prompt, and decode with M-RoPE position offsets."""
import pytest
import torch

from neuronx_distributed_inference_b200.utils.testing import perturb_constant_vectors  # noqa: E402

from neuronx_distributed_inference_b200.config import load_pretrained_config
from neuronx_distributed_inference_b200.utils.constants import get_model_cls


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


DEVICE = {"on_cpu": True, "dtype": "float32", "tol": 1.0}      # the GPU tier (test_multimodal_gpu.py) re-runs cases with cuda / bf16


def _build(model_type, hf, ckpt, **nc_kw):
    cls = get_model_cls(model_type, "image-text-to-text")
    nc = cls.get_neuron_config_cls()(batch_size=2, seq_len=64, max_context_length=32, torch_dtype=DEVICE["dtype"], on_cpu=DEVICE["on_cpu"],
                                     output_logits=True, **nc_kw)
    cfg = cls.get_config_cls()(nc, load_config=load_pretrained_config(ckpt))
    app = cls(ckpt, cfg)
    app.load(None, skip_warmup=True)
    return app


def test_qwen2_vl_matches_hf(tmp_path):
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration
    torch.manual_seed(0)
    cfg = Qwen2VLConfig(
        text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                         vocab_size=200, max_position_embeddings=512,
                         rope_parameters=dict(rope_type="default", mrope_section=[2, 3, 3], rope_theta=10000.0)),
        vision_config=dict(depth=2, embed_dim=32, hidden_size=64, num_heads=2, patch_size=4, spatial_merge_size=2,
                           temporal_patch_size=2, in_channels=3, mlp_ratio=2),
        image_token_id=150, video_token_id=151, vision_start_token_id=152, vision_end_token_id=153)
    hf = Qwen2VLForConditionalGeneration(cfg).eval()
    ckpt = str(tmp_path / "qwen2vl")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    app = _build("qwen2_vl", hf, ckpt)
    # two images: 4x4 and 4x8 patches -> 4 and 8 merged tokens
    grid = torch.tensor([[1, 4, 4], [1, 4, 8]])
    pix = torch.randn(16 + 32, 3 * 2 * 4 * 4)
    ids = torch.randint(1, 140, (2, 14))
    ids[0, 2:6] = 150
    ids[1, 1:9] = 150
    mask = torch.ones_like(ids)
    mask[0, 11:] = 0
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix, image_grid_thw=grid, mm_token_type_ids=(ids == 150).int())
        vis = hf.model.visual(pix, grid_thw=grid)
        vis = vis.pooler_output if hasattr(vis, "pooler_output") else vis
    got_vis = app.encode_images(pix, image_grid_thw=grid)
    assert _rel(got_vis, vis) < 1e-4 * DEVICE["tol"]
    if DEVICE["on_cpu"]:      # the reference's stand-alone image-encoding application name (wraps the same tower)
        from neuronx_distributed_inference_b200.models.qwen2_vl.modeling_qwen2_vl_vision import NeuronQwen2VLForImageEncoding
        enc = NeuronQwen2VLForImageEncoding(ckpt, app.config).load(None, skip_warmup=True)
        assert _rel(enc(pix, image_grid_thw=grid), vis) < 1e-4
    out = app(ids, attention_mask=mask, pixel_values=pix, image_grid_thw=grid)
    last = mask.sum(-1) - 1
    exp_last = exp.logits[torch.arange(2), last]
    assert _rel(out.logits[:, -1], exp_last) < 2e-4 * DEVICE["tol"]
    # one decode step (positions continue after the image-compressed rope index)
    nxt = exp_last.argmax(-1)
    ids2 = ids.clone()
    mask2 = torch.cat([mask, torch.zeros(2, 1, dtype=mask.dtype)], 1)
    ids2 = torch.cat([ids2, torch.zeros(2, 1, dtype=ids.dtype)], 1)
    ids2[torch.arange(2), last + 1] = nxt
    mask2[torch.arange(2), last + 1] = 1
    with torch.no_grad():
        exp2 = hf(input_ids=ids2, attention_mask=mask2, pixel_values=pix, image_grid_thw=grid,
                  mm_token_type_ids=(ids2 == 150).int()).logits[torch.arange(2), last + 1]
    out2 = app(nxt.view(2, 1), position_ids=(last + 1).view(2, 1).to(torch.int32))
    assert _rel(out2.logits[:, -1], exp2) < 2e-4 * DEVICE["tol"]


def test_pixtral_matches_hf(tmp_path):
    from transformers import LlavaConfig, LlavaForConditionalGeneration, MistralConfig, PixtralVisionConfig
    torch.manual_seed(0)
    cfg = LlavaConfig(
        vision_config=PixtralVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                                          patch_size=4, image_size=32, num_channels=3, head_dim=16).to_dict(),
        text_config=MistralConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                  num_key_value_heads=2, vocab_size=200, head_dim=16, sliding_window=None).to_dict(),
        image_token_index=150, projector_hidden_act="gelu", vision_feature_layer=-1, vision_feature_select_strategy="full")
    hf = LlavaForConditionalGeneration(cfg).eval()
    ckpt = str(tmp_path / "pixtral")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    app = _build("pixtral", hf, ckpt)
    pix = torch.randn(2, 3, 16, 16)
    sizes = torch.tensor([[8, 16], [16, 8]])        # 2x4 and 4x2 patches -> 8 tokens each
    ids = torch.randint(1, 140, (2, 14))
    ids[0, 1:9] = 150
    ids[1, 3:11] = 150
    mask = torch.ones_like(ids)
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix, image_sizes=sizes).logits
    out = app(ids, attention_mask=mask, pixel_values=pix, image_sizes=sizes)
    assert _rel(out.logits[:, -1], exp[:, -1]) < 2e-4 * DEVICE["tol"]


def test_qwen3_vl_matches_hf(tmp_path):
    from transformers import Qwen3VLConfig, Qwen3VLForConditionalGeneration
    torch.manual_seed(0)
    cfg = Qwen3VLConfig(
        text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
                         vocab_size=200, head_dim=16, max_position_embeddings=512,
                         rope_parameters=dict(rope_type="default", mrope_section=[4, 2, 2], mrope_interleaved=True, rope_theta=10000.0)),
        vision_config=dict(depth=3, hidden_size=32, intermediate_size=64, out_hidden_size=64, num_heads=2, patch_size=4,
                           spatial_merge_size=2, temporal_patch_size=2, in_channels=3, num_position_embeddings=64,
                           deepstack_visual_indexes=[0, 1]),
        image_token_id=150, video_token_id=151, vision_start_token_id=152, vision_end_token_id=153)
    hf = Qwen3VLForConditionalGeneration(cfg).eval()
    ckpt = str(tmp_path / "qwen3vl")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    app = _build("qwen3_vl", hf, ckpt)
    grid = torch.tensor([[1, 4, 4], [1, 4, 8]])
    pix = torch.randn(16 + 32, 3 * 2 * 4 * 4)
    ids = torch.randint(1, 140, (2, 14))
    ids[0, 2:6] = 150
    ids[1, 1:9] = 150
    mask = torch.ones_like(ids)
    mask[0, 11:] = 0
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix, image_grid_thw=grid, mm_token_type_ids=(ids == 150).int())
    out = app(ids, attention_mask=mask, pixel_values=pix, image_grid_thw=grid)
    last = mask.sum(-1) - 1
    assert _rel(out.logits[:, -1], exp.logits[torch.arange(2), last]) < 2e-4


def test_mllama_matches_hf(tmp_path):
    from transformers import MllamaConfig, MllamaForConditionalGeneration
    torch.manual_seed(0)
    cfg = MllamaConfig(
        vision_config=dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_global_layers=1, attention_heads=2,
                           patch_size=4, image_size=8, num_channels=3, max_num_tiles=2, vision_output_dim=96,
                           intermediate_layers_indices=[0, 1], supported_aspect_ratios=[[1, 1], [1, 2], [2, 1]]),
        text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
                         vocab_size=200, cross_attention_layers=[1], max_position_embeddings=512,
                         rope_parameters=dict(rope_type="default", rope_theta=10000.0), pad_token_id=0),
        image_token_index=150)
    hf = MllamaForConditionalGeneration(cfg)
    with torch.no_grad():        # gates are zero-initialised: open them so the cross-attention path matters
        for n, p in hf.named_parameters():
            if "gate" in n and p.numel() == 1:
                p.fill_(0.7)
    hf = hf.eval()
    ckpt = str(tmp_path / "mllama")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    app = _build("mllama", hf, ckpt)
    app.vision_model.intermediate_is_layer_output = True      # transformers 5.x oracle convention (see modeling_mllama.py)
    pix = torch.randn(2, 1, 2, 3, 8, 8)
    ar_ids = torch.tensor([[2], [1]])
    ar_mask = torch.tensor([[[1, 1]], [[1, 0]]])
    ids = torch.randint(1, 140, (2, 10))
    ids[:, 1] = 150
    mask = torch.ones_like(ids)
    mask[1, 8:] = 0
    cam = torch.zeros(2, 10, 1, 2, dtype=torch.long)
    cam[0, 1:, 0, :] = 1
    cam[1, 1:, 0, 0] = 1
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix, aspect_ratio_ids=ar_ids, aspect_ratio_mask=ar_mask,
                 cross_attention_mask=cam).logits
    out = app(ids, attention_mask=mask, pixel_values=pix, aspect_ratio_ids=ar_ids, aspect_ratio_mask=ar_mask,
              cross_attention_mask=cam)
    last = mask.sum(-1) - 1
    assert _rel(out.logits[:, -1], exp[torch.arange(2), last]) < 2e-4
    # decode one token: cross K/V come from the per-line buffers
    nxt = exp[torch.arange(2), last].argmax(-1)
    ids2 = torch.cat([ids, torch.zeros(2, 1, dtype=ids.dtype)], 1)
    mask2 = torch.cat([mask, torch.zeros(2, 1, dtype=mask.dtype)], 1)
    ids2[torch.arange(2), last + 1] = nxt
    mask2[torch.arange(2), last + 1] = 1
    cam2 = torch.cat([cam, cam[:, -1:]], 1)
    cam2[1, 8] = cam[1, 7]
    with torch.no_grad():
        exp2 = hf(input_ids=ids2, attention_mask=mask2, pixel_values=pix, aspect_ratio_ids=ar_ids, aspect_ratio_mask=ar_mask,
                  cross_attention_mask=cam2).logits[torch.arange(2), last + 1]
    out2 = app(nxt.view(2, 1), position_ids=(last + 1).view(2, 1).to(torch.int32))
    assert _rel(out2.logits[:, -1], exp2) < 2e-4


def test_llama4_multimodal_matches_hf(tmp_path):
    from transformers import Llama4Config, Llama4ForConditionalGeneration
    torch.manual_seed(0)
    cfg = Llama4Config(
        vision_config=dict(hidden_size=32, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, patch_size=4,
                           image_size=16, num_channels=3, pixel_shuffle_ratio=0.5, projector_input_dim=48, projector_output_dim=48,
                           vision_output_dim=48, rope_parameters=dict(rope_theta=10000.0, rope_type="default")),
        text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                         vocab_size=200, head_dim=16, num_local_experts=4, num_experts_per_tok=1, intermediate_size_mlp=128,
                         interleave_moe_layer_step=2, attention_chunk_size=8, no_rope_layers=[1, 0], use_qk_norm=True,
                         max_position_embeddings=256),
        image_token_index=150)
    hf = Llama4ForConditionalGeneration(cfg).eval()
    ckpt = str(tmp_path / "llama4mm")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    app = _build("llama4", hf, ckpt)
    pix = torch.randn(2, 3, 16, 16)           # two tiles -> 4 tokens each after the 0.5 pixel shuffle
    ids = torch.randint(1, 140, (2, 12))
    ids[0, 1:5] = 150
    ids[1, 2:6] = 150
    mask = torch.ones_like(ids)
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix).logits
    out = app(ids, attention_mask=mask, pixel_values=pix)
    assert _rel(out.logits[:, -1], exp[:, -1]) < 2e-4


def test_llava_1_5_matches_hf(tmp_path):
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig, LlavaForConditionalGeneration
    from neuronx_distributed_inference_b200.contrib.models.llava import NeuronLlavaForCausalLM
    torch.manual_seed(0)
    cfg = LlavaConfig(
        vision_config=CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2, patch_size=4,
                                       image_size=16, projection_dim=16).to_dict(),
        text_config=LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                vocab_size=200).to_dict(),
        image_token_index=150, vision_feature_layer=-2, vision_feature_select_strategy="default")
    hf = LlavaForConditionalGeneration(cfg).eval()
    ckpt = str(tmp_path / "llava")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    cls = NeuronLlavaForCausalLM
    nc = cls.get_neuron_config_cls()(batch_size=2, seq_len=64, max_context_length=32, torch_dtype="float32", on_cpu=True, output_logits=True)
    app = cls(ckpt, cls.get_config_cls()(nc, load_config=load_pretrained_config(ckpt)))
    app.load(None, skip_warmup=True)
    pix = torch.randn(2, 3, 16, 16)                      # 16 patches each
    ids = torch.randint(1, 140, (2, 22))
    ids[0, 1:17] = 150
    ids[1, 4:20] = 150
    mask = torch.ones_like(ids)
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix).logits
    out = app(ids, attention_mask=mask, pixel_values=pix)
    assert _rel(out.logits[:, -1], exp[:, -1]) < 2e-4


def test_qwen2_5_vl_matches_hf(tmp_path):
    from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration
    from neuronx_distributed_inference_b200.contrib.models.qwen2_5_vl import NeuronQwen25VLForCausalLM
    torch.manual_seed(0)
    cfg = Qwen2_5_VLConfig(
        text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                         vocab_size=200, max_position_embeddings=512,
                         rope_parameters=dict(rope_type="default", mrope_section=[2, 3, 3], rope_theta=10000.0)),
        vision_config=dict(depth=3, hidden_size=32, intermediate_size=64, out_hidden_size=64, num_heads=2, patch_size=4, spatial_merge_size=2,
                           temporal_patch_size=2, in_channels=3, window_size=16, fullatt_block_indexes=[1],
                           tokens_per_second=1),      # transformers 5.5 multiplies an IMAGE's start position by this interval;
        # the released model uses t = start for images (interval only spaces video frames): 1 makes both agree
        image_token_id=150, video_token_id=151, vision_start_token_id=152, vision_end_token_id=153)
    hf = Qwen2_5_VLForConditionalGeneration(cfg).eval()
    ckpt = str(tmp_path / "qwen25vl")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    cls = NeuronQwen25VLForCausalLM
    nc = cls.get_neuron_config_cls()(batch_size=2, seq_len=64, max_context_length=32, torch_dtype="float32", on_cpu=True, output_logits=True)
    app = cls(ckpt, cls.get_config_cls()(nc, load_config=load_pretrained_config(ckpt)))
    app.load(None, skip_warmup=True)
    grid = torch.tensor([[1, 4, 8], [1, 8, 4]])             # 8 merged tokens each; 2x2-unit windows
    pix = torch.randn(64, 3 * 2 * 4 * 4)
    ids = torch.randint(1, 140, (2, 14))
    ids[0, 2:10] = 150
    ids[1, 1:9] = 150
    mask = torch.ones_like(ids)
    with torch.no_grad():
        vis = hf.model.visual(pix, grid_thw=grid)
        vis = vis.pooler_output if hasattr(vis, "pooler_output") else vis
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix, image_grid_thw=grid, mm_token_type_ids=(ids == 150).int())
    assert _rel(app.encode_images(pix, image_grid_thw=grid), vis) < 1e-4
    out = app(ids, attention_mask=mask, pixel_values=pix, image_grid_thw=grid)
    assert _rel(out.logits[:, -1], exp.logits[:, -1]) < 2e-4


def test_mistral3_matches_hf(tmp_path):
    """Mistral-Small-3.1: Pixtral tower + RMSNorm / 2x2 patch-merger projector + Mistral decoder."""
    from transformers import Mistral3Config, Mistral3ForConditionalGeneration, MistralConfig, PixtralVisionConfig
    torch.manual_seed(0)
    cfg = Mistral3Config(
        vision_config=PixtralVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2,
                                          patch_size=4, image_size=32, num_channels=3, head_dim=16).to_dict(),
        text_config=MistralConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                  num_key_value_heads=2, vocab_size=200, head_dim=16, sliding_window=None).to_dict(),
        image_token_index=150, projector_hidden_act="gelu", vision_feature_layer=-1, spatial_merge_size=2)
    hf = Mistral3ForConditionalGeneration(cfg).eval()
    ckpt = str(tmp_path / "mistral3")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    app = _build("mistral3", hf, ckpt)
    pix = torch.randn(2, 3, 16, 16)
    sizes = torch.tensor([[8, 16], [16, 8]])        # 2x4 and 4x2 patches -> merged 1x2 and 2x1 -> 2 tokens each
    ids = torch.randint(1, 140, (2, 14))
    ids[0, 1:3] = 150
    ids[1, 5:7] = 150
    mask = torch.ones_like(ids)
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix, image_sizes=sizes).logits
    out = app(ids, attention_mask=mask, pixel_values=pix, image_sizes=sizes)
    assert _rel(out.logits[:, -1], exp[:, -1]) < 2e-4


def test_gemma3_vision_matches_hf(tmp_path):
    """SigLIP tower + pooled projector + Gemma-3 decoder with bidirectional attention inside each image's soft tokens (incl. the
    sliding-window layers), then a decode step."""
    from transformers import Gemma3Config, Gemma3ForConditionalGeneration
    torch.manual_seed(0)
    cfg = Gemma3Config(
        text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
                         vocab_size=200, head_dim=16, sliding_window=4, query_pre_attn_scalar=16, max_position_embeddings=256,
                         layer_types=["sliding_attention", "sliding_attention", "full_attention"]),
        vision_config=dict(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2, image_size=16, patch_size=4,
                           num_channels=3),
        mm_tokens_per_image=4, image_token_index=150, boi_token_index=151, eoi_token_index=152)
    hf = Gemma3ForConditionalGeneration(cfg).eval()
    hf.model.multi_modal_projector.mm_input_projection_weight.data.normal_(0, 0.1)      # HF initialises the projector to zeros
    hf.model.multi_modal_projector.mm_soft_emb_norm.weight.data.normal_(0, 0.1)
    ckpt = str(tmp_path / "gemma3v")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    app = _build("gemma3", hf, ckpt)
    pix = torch.randn(3, 3, 16, 16)
    ids = torch.randint(1, 140, (2, 16))
    ids[0, 2:6] = 150                      # row 0: one image
    ids[1, 1:5] = 150                      # row 1: two images separated by a text token
    ids[1, 6:10] = 150
    mask = torch.ones_like(ids)
    tt = (ids == 150).int()
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix, token_type_ids=tt).logits
        feats = hf.model.get_image_features(pix)
        feats = feats.pooler_output if hasattr(feats, "pooler_output") else feats
    assert _rel(app.encode_images(pix), torch.cat([f for f in feats]).reshape(-1, 64)) < 1e-4
    out = app(ids, attention_mask=mask, pixel_values=pix)
    assert _rel(out.logits[:, -1], exp[:, -1]) < 2e-4
    nxt = exp[:, -1].argmax(-1)
    ids2 = torch.cat([ids, nxt.view(2, 1)], 1)
    with torch.no_grad():
        exp2 = hf(input_ids=ids2, attention_mask=torch.ones_like(ids2), pixel_values=pix, token_type_ids=(ids2 == 150).int()).logits[:, -1]
    out2 = app(nxt.view(2, 1), position_ids=torch.full((2, 1), 16, dtype=torch.int32))
    assert _rel(out2.logits[:, -1], exp2) < 2e-4


def test_mllama_image_tiling_and_vision_mask_helpers_match_transformers():
    """reference models/mllama/{image_transform,utils}.py — canvas choice, fit-to-canvas size, tiling, aspect-ratio ids / masks and
    the image -> text-range visibility are checked against the Hugging Face Mllama processor functions."""
    import random
    from transformers.models.mllama import image_processing_mllama as I
    from transformers.models.mllama.processing_mllama import get_cross_attention_token_mask
    from neuronx_distributed_inference_b200.models.mllama import utils as U
    from neuronx_distributed_inference_b200.models.mllama.image_transform import VariableSizeImageTransform, custom_image_preprocessing
    tf = VariableSizeImageTransform(size=56, max_num_tiles=4)
    assert U.get_all_supported_aspect_ratios(4) == I.get_all_supported_aspect_ratios(4)
    rng = random.Random(0)
    for _ in range(200):
        h, w = rng.randint(8, 400), rng.randint(8, 400)
        canvas = I.get_optimal_tiled_canvas(h, w, 4, 56)
        assert tf.get_optimal_tiled_canvas(h, w) == tuple(canvas)
        assert tf.get_image_size_fit_to_canvas(h, w, *canvas) == tuple(I.get_image_size_fit_to_canvas(h, w, canvas[0], canvas[1], 56))
    img = torch.rand(3, 50, 100)
    tiles, ar = tf(img)
    assert ar == (1, 2) and tiles.shape == (2, 3, 56, 56)
    exp = I.split_to_tiles(torch.arange(3 * 112 * 112.).view(1, 3, 112, 112), 2, 2)[0]
    assert torch.equal(tf.split_to_tiles(torch.arange(3 * 112 * 112.).view(3, 112, 112), 2, 2), exp)
    ars = [[(1, 2), (2, 2)], [(1, 1)]]
    assert torch.equal(U.convert_aspect_ratios_to_ids(ars, 4), torch.as_tensor(I.convert_aspect_ratios_to_ids(ars, 4)))
    assert torch.equal(U.get_aspect_ratio_mask(ars, 4), torch.as_tensor(I.build_aspect_ratio_mask(ars, 4)))
    for ids in ([5, 9, 1, 2, 9, 9, 3, 4], [1, 2, 3], [9, 1, 1], [1, 9, 9], [9, 9, 9, 1]):
        assert U.create_vision_mask(ids, 9) == get_cross_attention_token_mask(ids, 9)
    dense = U.vision_mask_to_dense([U.create_vision_mask([5, 9, 1, 2, 9, 3], 9)], [[2, 4]], 6, 2, 4)
    assert dense[0, :, 0].sum(-1).tolist() == [0, 2, 2, 2, 0, 0] and dense[0, :, 1].sum(-1).tolist() == [0, 0, 0, 0, 4, 4]
    pix, ids, mask, counts = custom_image_preprocessing([[torch.rand(3, 60, 200)], [torch.rand(3, 50, 50), torch.rand(3, 120, 60)]], 56, 4)
    assert pix.shape == (2, 2, 4, 3, 56, 56) and counts.tolist() == [[4, 0], [1, 2]] and ids.shape == (2, 2) and mask.shape == (2, 2, 4)
    assert "<|image|>" in U.add_instruct("hi", True) and "<|image|>" not in U.add_instruct("hi", False)


@pytest.mark.parametrize("resampler", [False, True])
def test_idefics_matches_hf(resampler, tmp_path):
    """Gated cross-attention decoder + CLIP tower: prefill with two images per row (the second hidden from part of the text, one row
    with text that sees no image at all) and two decode steps, against Hugging Face."""
    from transformers import IdeficsConfig, IdeficsForVisionText2Text
    torch.manual_seed(0)
    cfg = IdeficsConfig(vocab_size=160, additional_vocab_size=4, hidden_size=64, intermediate_size=128, num_hidden_layers=4, num_attention_heads=4,
                        cross_layer_interval=2, qk_layer_norms=True, alpha_type="vector", alpha_initializer="normal", alphas_initializer_range=0.5,
                        vision_config=dict(embed_dim=32, image_size=16, patch_size=4, num_hidden_layers=2, num_attention_heads=2, intermediate_size=64),
                        perceiver_config=dict(use_resampler=resampler, resampler_n_latents=5, resampler_depth=2, resampler_n_heads=2,
                                              resampler_head_dim=8, qk_layer_norms_perceiver=True),
                        use_resampler=resampler, tie_word_embeddings=False, freeze_text_layers=False, freeze_vision_layers=False,
                        max_position_embeddings=128)
    hf = IdeficsForVisionText2Text(cfg).eval()
    for n, p in hf.named_parameters():
        if "layer_norm" in n or "layernorm" in n:
            p.data.add_(torch.randn_like(p) * 0.1)
    ckpt = str(tmp_path / "idefics")
    perturb_constant_vectors(hf)
    hf.save_pretrained(ckpt)
    app = _build("idefics", hf, ckpt)
    ids = torch.randint(1, 164, (2, 12))                     # includes ids from the additional vocabulary
    mask = torch.ones_like(ids)
    pix = torch.randn(2, 2, 3, 16, 16)
    iam = torch.zeros(2, 12, 2, dtype=torch.long)
    iam[0, 2:, 0] = 1
    iam[0, 7:, 1] = 1
    iam[1, 5:, 0] = 1                                        # row 1: the first five tokens see no image, image 1 is never used
    with torch.no_grad():
        exp = hf(input_ids=ids, attention_mask=mask, pixel_values=pix, image_attention_mask=iam)
        feat = hf.model.vision_model(pixel_values=pix.flatten(0, 1)).last_hidden_state
        if resampler:
            feat = hf.model.perceiver_resampler(feat)
    assert _rel(app.encode_images(pix), feat.reshape(2, -1, 32)) < 1e-4 and feat.shape[1] == (5 if resampler else 17)
    out = app(ids, attention_mask=mask, pixel_values=pix, image_attention_mask=iam)
    assert _rel(out.logits[:, -1], exp.logits[:, -1]) < 2e-4
    past, tok, last_iam = exp.past_key_values, exp.logits[:, -1].argmax(-1), iam[:, -1:]
    for step in range(2):
        with torch.no_grad():
            e = hf(input_ids=tok.view(2, 1), past_key_values=past, pixel_values=pix, image_attention_mask=last_iam,
                   attention_mask=torch.ones(2, 13 + step, dtype=torch.long))
        o = app(tok.view(2, 1), position_ids=torch.full((2, 1), 12 + step, dtype=torch.int32))
        assert _rel(o.logits[:, -1], e.logits[:, -1]) < 2e-4, step
        past, tok = e.past_key_values, e.logits[:, -1].argmax(-1)
