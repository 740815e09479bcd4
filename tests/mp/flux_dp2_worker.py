"""gloo worker (2 ranks, tp=1): FLUX context-parallel backbone and CFG-parallel pipeline == the single-replica computation."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

BACKBONE = dict(num_layers=2, num_single_layers=2, attention_head_dim=16, num_attention_heads=2, in_channels=16, joint_attention_dim=24,
                pooled_projection_dim=20, axes_dims_rope=(4, 6, 6), guidance_embeds=True)
CLIP = dict(vocab_size=100, hidden_size=20, intermediate_size=40, num_hidden_layers=2, num_attention_heads=2, max_position_embeddings=16,
            hidden_act="quick_gelu", layer_norm_eps=1e-5, eos_token_id=2)
T5 = dict(vocab_size=100, d_model=24, d_kv=8, d_ff=48, num_layers=2, num_heads=3, relative_attention_num_buckets=8,
          relative_attention_max_distance=16, layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu")
VAE = dict(latent_channels=4, out_channels=3, block_out_channels=(8, 16), layers_per_block=1, norm_num_groups=4, scaling_factor=0.36, shift_factor=0.11)


def main():
    import torch.distributed as dist
    from neuronx_distributed_inference_b200.config import NeuronConfig
    from neuronx_distributed_inference_b200.models.diffusers.flux.application import NeuronFluxApplication, get_flux_parallelism_config
    mode = sys.argv[1]
    assert get_flux_parallelism_config(1, mode == "cp", mode == "cfg") == int(os.environ["WORLD_SIZE"])
    nc = NeuronConfig(batch_size=1, torch_dtype="float32", on_cpu=True, tp_degree=1)
    app = NeuronFluxApplication(None, nc, BACKBONE, CLIP, T5, VAE, height=32, width=32, context_parallel_enabled=mode == "cp",
                                cfg_parallel_enabled=mode == "cfg").load(random_weights=True, seed=0)
    g = torch.Generator().manual_seed(3)
    ids1, ids2, neg = torch.randint(3, 99, (1, 8), generator=g), torch.randint(3, 99, (1, 8), generator=g), torch.randint(3, 99, (1, 6), generator=g)
    lat = torch.randn(1, 4, 16, 16, generator=g)
    kw = dict(num_inference_steps=2, latents=lat.clone(), output_type="latent")
    if mode == "cfg":
        kw.update(negative_t5_input_ids=neg, true_cfg_scale=3.0)
    got = app(ids1, ids2, **kw)
    # single-replica reference on this rank: same weights (same seed), parallel mode off
    app.transformer.cp_group = None
    app.pipe.cfg_group = None
    kw["latents"] = lat.clone()
    ref = app(ids1, ids2, **kw)
    ok = torch.allclose(got, ref, atol=1e-5, rtol=1e-4) and bool(torch.isfinite(got).all())
    if mode == "cfg":
        kw.pop("negative_t5_input_ids")
        kw["latents"] = lat.clone()
        ok = ok and not torch.equal(app(ids1, ids2, **kw), ref)                     # the negative branch really enters the result
    flags = [torch.zeros(1) for _ in range(2)]
    dist.all_gather(flags, torch.tensor([1.0 if ok else 0.0]))
    if dist.get_rank() == 0:
        print('{"ok": %s}' % ("true" if all(f.item() == 1.0 for f in flags) else "false"), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
