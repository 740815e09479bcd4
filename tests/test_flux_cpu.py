"""FLUX stack: CLIP / T5 encoders vs transformers, MMDiT backbone vs an independent functional re-statement of the
architecture, VAE decoder shapes, scheduler + packing algebra and an end-to-end tiny pipeline."""
import math

import torch

from neuronx_distributed_inference_b200.config import NeuronConfig
from neuronx_distributed_inference_b200.models.diffusers.flux.application import NeuronFluxApplication
from neuronx_distributed_inference_b200.models.diffusers.flux.pipeline import (FlowMatchEulerScheduler, pack_latents, unpack_latents)
from neuronx_distributed_inference_b200.modules.checkpoint import load_sharded

BACKBONE = dict(num_layers=2, num_single_layers=2, attention_head_dim=16, num_attention_heads=2, in_channels=16,
                joint_attention_dim=24, pooled_projection_dim=20, axes_dims_rope=(4, 6, 6), guidance_embeds=True)
CLIP = dict(vocab_size=100, hidden_size=20, intermediate_size=40, num_hidden_layers=2, num_attention_heads=2, max_position_embeddings=16,
            hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2)
T5 = dict(vocab_size=100, d_model=24, d_kv=8, d_ff=48, num_layers=2, num_heads=3, relative_attention_num_buckets=8,
          relative_attention_max_distance=16, layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu")
VAE = dict(latent_channels=4, out_channels=3, block_out_channels=(8, 16), layers_per_block=1, norm_num_groups=4, scaling_factor=0.36,
           shift_factor=0.11)


def _app():
    nc = NeuronConfig(batch_size=1, torch_dtype="float32", on_cpu=True)
    return NeuronFluxApplication(None, nc, BACKBONE, CLIP, T5, VAE, height=32, width=32).load(random_weights=True)


def test_text_encoders_match_transformers():
    from transformers import CLIPTextConfig, CLIPTextModel, T5Config, T5EncoderModel
    from neuronx_distributed_inference_b200.models.diffusers.flux.clip.modeling_clip import convert_clip_state_dict
    from neuronx_distributed_inference_b200.models.diffusers.flux.t5.modeling_t5 import convert_t5_state_dict
    app = _app()
    torch.manual_seed(0)
    hf_clip = CLIPTextModel(CLIPTextConfig(**{**CLIP, "bos_token_id": 0, "pad_token_id": 1})).eval()
    load_sharded(app.clip, convert_clip_state_dict(hf_clip.state_dict(), app.clip_config), torch.float32, strict=False)
    ids = torch.randint(3, 99, (2, 9))
    ids[:, -1] = 99
    with torch.no_grad():
        exp = hf_clip(ids)
    last, pooled = app.clip(ids)
    assert torch.allclose(last, exp.last_hidden_state, atol=1e-5) and torch.allclose(pooled, exp.pooler_output, atol=1e-5)
    hf_t5 = T5EncoderModel(T5Config(**T5)).eval()
    load_sharded(app.t5, convert_t5_state_dict(hf_t5.state_dict(), app.t5_config), torch.float32, strict=False)
    with torch.no_grad():
        exp = hf_t5(ids).last_hidden_state
    assert torch.allclose(app.t5(ids), exp, atol=1e-5)


def _ref_backbone(m, x, ctx, pooled, t, img_ids, txt_ids, g):
    """Functional re-statement of the MMDiT forward from the architecture description, reading the module's weights."""
    from neuronx_distributed_inference_b200.models.diffusers.embeddings import timestep_sinusoid
    F = torch.nn.functional
    H, D = 2, 16
    lin = lambda l, v: F.linear(v, l.weight, l.bias)  # noqa: E731

    def emb(e, v):
        return lin(e.linear_2, F.silu(lin(e.linear_1, v)))
    tt = m.time_text_embed
    temb = emb(tt.timestep_embedder, timestep_sinusoid(t * 1000)) + emb(tt.guidance_embedder, timestep_sinusoid(g * 1000)) + emb(tt.text_embedder, pooled)
    ids = torch.cat([txt_ids, img_ids])
    ang = torch.cat([ids[:, i:i + 1].double() * (1.0 / (10000.0 ** (torch.arange(0, d, 2).double() / d)))[None] for i, d in enumerate((4, 6, 6))], -1)
    rot = torch.polar(torch.ones_like(ang), ang)

    def rope(v):       # v [B,N,H,D]
        vc = torch.view_as_complex(v.double().reshape(*v.shape[:-1], -1, 2))
        return torch.view_as_real(vc * rot[None, :, None]).flatten(3).float()

    def rms(v, w):
        return v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + 1e-6) * w

    def qkv(a, v):
        B, N, _ = v.shape
        q, k, vv = F.linear(v, a.proj.weight, a.proj.bias).view(B, N, 3, H, D).unbind(2)
        return rms(q, a.norm_q.weight), rms(k, a.norm_k.weight), vv

    def attn(q, k, v):
        o = F.scaled_dot_product_attention(rope(q).transpose(1, 2), rope(k).transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2).flatten(2)
    ln = lambda v: F.layer_norm(v, (32,), eps=1e-6)  # noqa: E731
    x, c = lin(m.x_embedder, x), lin(m.context_embedder, ctx)
    for b in m.transformer_blocks:
        m1 = lin(b.norm1.linear, F.silu(temb))[:, None].chunk(6, -1)
        c1 = lin(b.norm1_context.linear, F.silu(temb))[:, None].chunk(6, -1)
        q, k, v = qkv(b.attn, ln(x) * (1 + m1[1]) + m1[0])
        cq, ck, cv = qkv(b.attn_context, ln(c) * (1 + c1[1]) + c1[0])
        o = attn(torch.cat([cq, q], 1), torch.cat([ck, k], 1), torch.cat([cv, v], 1))
        nc_ = c.shape[1]
        x = x + m1[2] * lin(b.to_out, o[:, nc_:])
        c = c + c1[2] * lin(b.to_add_out, o[:, :nc_])
        ff = lambda f, v: lin(f.fc2, F.gelu(lin(f.fc1, v), approximate="tanh"))  # noqa: E731
        x = x + m1[5] * ff(b.ff, ln(x) * (1 + m1[4]) + m1[3])
        c = c + c1[5] * ff(b.ff_context, ln(c) * (1 + c1[4]) + c1[3])
    h = torch.cat([c, x], 1)
    for b in m.single_transformer_blocks:
        sh, sc, gate = lin(b.norm.linear, F.silu(temb))[:, None].chunk(3, -1)
        hn = ln(h) * (1 + sc) + sh
        q, k, v = qkv(b.attn, hn)
        o = attn(q, k, v)
        mlp = F.gelu(lin(b.proj_mlp, hn), approximate="tanh")
        h = h + gate * (lin(b.proj_out_attn, o) + F.linear(mlp, b.proj_out_mlp.weight))
    x = h[:, c.shape[1]:]
    sc, sh = lin(m.norm_out.linear, F.silu(temb))[:, None].chunk(2, -1)
    return lin(m.proj_out, ln(x) * (1 + sc) + sh)


def test_backbone_matches_functional_reference():
    app = _app()
    torch.manual_seed(1)
    x, ctx, pooled = torch.randn(1, 16, 16), torch.randn(1, 5, 24), torch.randn(1, 20)
    t, g = torch.tensor([0.7]), torch.tensor([3.5])
    img_ids = torch.zeros(16, 3)
    img_ids[:, 1] = torch.arange(16) // 4
    img_ids[:, 2] = torch.arange(16) % 4
    txt_ids = torch.zeros(5, 3)
    got = app.transformer(x, ctx, pooled, t, img_ids, txt_ids, g)
    exp = _ref_backbone(app.transformer, x, ctx, pooled, t, img_ids, txt_ids, g)
    assert got.shape == (1, 16, 16) and torch.allclose(got, exp, atol=2e-5), (got - exp).abs().max()


def test_scheduler_packing_and_pipeline():
    lat = torch.randn(2, 4, 8, 6)
    assert torch.equal(unpack_latents(pack_latents(lat), 8, 6), lat)
    s = FlowMatchEulerScheduler()
    ts = s.set_timesteps(4, mu=0.8)
    assert ts.shape == (4,) and float(s.sigmas[-1]) == 0.0 and bool((s.sigmas[:-1] > s.sigmas[1:]).all())
    # exact integration of a constant velocity field: x_T + (0 - sigma_0) * v
    x = torch.zeros(1, 3)
    for i in range(4):
        x = s.step(torch.ones(1, 3), i, x)
    assert torch.allclose(x, -s.sigmas[0] * torch.ones(1, 3), atol=1e-6)
    app = _app()
    img = app(torch.randint(3, 99, (1, 8)), torch.randint(3, 99, (1, 6)), num_inference_steps=2,
              generator=torch.Generator().manual_seed(0))
    assert img.shape == (1, 3, 32, 32) and torch.isfinite(img).all() and 0.0 <= float(img.min()) and float(img.max()) <= 1.0


def test_flux_control_and_fill_pipelines():
    """Image-conditioned FLUX (reference NeuronFluxControlPipeline / NeuronFluxFillPipeline): the backbone sees
    [latents | control latents] (2 x 64 -> here 2 x 16 channels) resp. [latents | masked-image latents | mask footprint]."""
    nc = NeuronConfig(batch_size=1, torch_dtype="float32", on_cpu=True)
    lat_c, f = VAE["latent_channels"], 2                                         # two VAE resolutions -> scale factor 2
    for task, extra_ch in (("control", lat_c * 4), ("fill", lat_c * 4 + f * f * 4)):
        app = NeuronFluxApplication(None, nc, {**BACKBONE, "in_channels": lat_c * 4 + extra_ch, "out_channels": lat_c * 4}, CLIP, T5, VAE,
                                    height=16, width=16, task=task).load(random_weights=True)
        seen = []
        inner = app.pipe.transformer
        app.pipe.transformer = lambda x, *a: (seen.append(x.clone()), inner(x, *a))[1]
        ids1, ids2 = torch.randint(3, 99, (1, 8)), torch.randint(3, 99, (1, 8))
        img = torch.rand(1, 3, 16, 16)
        g = torch.Generator().manual_seed(0)
        if task == "control":
            out = app(ids1, ids2, num_inference_steps=2, generator=g, control_image=img)
            cond = app.pipe.conditioning(1, 8, 8, control_image=img)
            from neuronx_distributed_inference_b200.models.diffusers.flux.pipeline import pack_latents
            assert torch.allclose(cond, pack_latents(app.vae_encoder(img * 2 - 1)), atol=1e-6)
        else:
            mask = torch.zeros(1, 1, 16, 16)
            mask[..., 4:12, 4:12] = 1
            out = app(ids1, ids2, num_inference_steps=2, generator=g, image=img, mask_image=mask)
            cond = app.pipe.conditioning(1, 8, 8, image=img, mask_image=mask)
            mk = cond[0, :, lat_c * 4:]                                            # [tokens, f*f*4] mask footprint per 2x2 latent patch
            assert set(mk.unique().tolist()) <= {0.0, 1.0} and mk.shape == (16, f * f * 4)
            tok = mk.view(4, 4, -1)                                                # token grid 4x4, each covers 4x4 pixels
            assert tok[1:3, 1:3].min() == 1 and tok[0].max() == 0 and tok[:, 0].max() == 0
        assert out.shape == (1, 3, 16, 16) and torch.isfinite(out).all()
        assert len(seen) == 2 and seen[0].shape == (1, 16, lat_c * 4 + extra_ch)
        assert torch.equal(seen[0][..., lat_c * 4:], seen[1][..., lat_c * 4:]) and torch.allclose(seen[0][..., lat_c * 4:], cond)
        assert not torch.equal(seen[0][..., : lat_c * 4], seen[1][..., : lat_c * 4])       # the latents moved, the conditioning did not
    import pytest
    with pytest.raises(ValueError):
        app.pipe.conditioning(1, 8, 8, image=img)


def test_vae_encoder_shapes_and_state_dict_names():
    from neuronx_distributed_inference_b200.models.diffusers.flux.vae.modeling_vae import NeuronVAEEncoder, convert_vae_encoder_state_dict
    app = _app()
    enc = NeuronVAEEncoder(app.vae_config)
    z = enc(torch.rand(2, 3, 32, 32) * 2 - 1)
    assert z.shape == (2, VAE["latent_channels"], 16, 16)
    # diffusers AutoencoderKL names -> ours (every parameter of the module must be reachable)
    names = {}
    for k in enc.state_dict():
        d = (k.replace("mid_res1.", "mid_block.resnets.0.").replace("mid_res2.", "mid_block.resnets.1.").replace("mid_attn.", "mid_block.attentions.0.")
             .replace(".downsample.", ".downsamplers.0.conv.").replace(".to_out.", ".to_out.0."))
        names["encoder." + d] = k
    sd = convert_vae_encoder_state_dict({d: torch.zeros(1) for d in names} | {"decoder.conv_in.weight": torch.zeros(1)})
    assert set(sd) == set(enc.state_dict())
