"""FLUX.1 MMDiT backbone (reference models/diffusers/flux/modeling_flux.py, ≈1500 LoC): ``num_layers`` double-stream blocks
(image and text tokens keep separate weights, attend jointly) followed by ``num_single_layers`` single-stream blocks (fused
attention + MLP), adaLN-Zero modulation from (timestep, guidance, pooled text), 3-axis rotary embedding.

Tensor parallel over attention heads / MLP columns through the engine's parallel layers; parameter names follow the
diffusers checkpoint layout so ``transformer/diffusion_pytorch_model*.safetensors`` load after q/k/v fusion."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from .... import ops
from ....config import InferenceConfig, NeuronConfig
from ....modules.gqa import GroupQueryAttention_O, GroupQueryAttention_QKV
from ....modules.norm import RMSNorm
from ....parallel import mappings
from ....parallel.layers import ColumnParallelLinear, RowParallelLinear
from ....parallel.state import get_tensor_model_parallel_group
from ..embeddings import CombinedTimestepGuidanceTextProjEmbeddings, rope_nd
from ..normalization import AdaLayerNormContinuous, AdaLayerNormZero


class FluxBackboneInferenceConfig(InferenceConfig):
    def get_required_attributes(self) -> List[str]:
        return ["num_layers", "num_single_layers", "attention_head_dim", "num_attention_heads", "in_channels",
                "joint_attention_dim", "pooled_projection_dim"]

    def add_derived_config(self):
        self.num_cores_per_group = 1
        if not hasattr(self, "axes_dims_rope"):
            self.axes_dims_rope = (16, 56, 56)
        if not hasattr(self, "guidance_embeds"):
            self.guidance_embeds = True
        if not hasattr(self, "patch_size"):
            self.patch_size = 1
        if not hasattr(self, "out_channels") or self.out_channels is None:
            self.out_channels = self.in_channels


def _rope(x, cos, sin):
    return ops.apply_rope(x, cos, sin, True)


def _attend(q, k, v, scale, cp=None):
    """Joint bidirectional attention over [text | image] tokens: the flash kernel's non-causal mode on CUDA.
    ``cp = (group, n_txt)``: context parallel — this rank holds all text tokens and ITS slice of the image tokens; keys / values of
    the other slices are all-gathered (already rotated with their own positions), queries stay local."""
    if cp is not None:
        g, n = cp
        k = torch.cat([k[:, :n], mappings.all_gather(k[:, n:].contiguous(), 1, g)], 1)
        v = torch.cat([v[:, :n], mappings.all_gather(v[:, n:].contiguous(), 1, g)], 1)
    return ops.attention_prefill(q.contiguous(), k.contiguous(), v.contiguous(), scale, causal=False)


class _QKV(nn.Module):
    """fused q/k/v projection + per-head RMSNorm of q and k."""

    def __init__(self, dim, heads, hd, dtype, device):
        super().__init__()
        self.proj = GroupQueryAttention_QKV(dim, hd, heads, heads, None, dtype, True, None, device)
        self.norm_q = RMSNorm(hd, 1e-6, dtype, device=device)
        self.norm_k = RMSNorm(hd, 1e-6, dtype, device=device)
        self.h, self.hd = self.proj.n_q, hd

    def forward(self, x):
        B, N, _ = x.shape
        q, k, v = self.proj(x).view(B, N, 3 * self.h, self.hd).split(self.h, 2)
        return self.norm_q(q), self.norm_k(k), v


class _FF(nn.Module):
    def __init__(self, dim, dtype, device):
        super().__init__()
        self.fc1 = ColumnParallelLinear(dim, 4 * dim, bias=True, gather_output=False, dtype=dtype, device=device)
        self.fc2 = RowParallelLinear(4 * dim, dim, bias=True, input_is_parallel=True, dtype=dtype, device=device)

    def forward(self, x):
        return self.fc2(nn.functional.gelu(self.fc1(x), approximate="tanh"))


class FluxDoubleBlock(nn.Module):
    def __init__(self, dim, heads, hd, dtype, device):
        super().__init__()
        self.norm1 = AdaLayerNormZero(dim, 6, dtype, device)
        self.norm1_context = AdaLayerNormZero(dim, 6, dtype, device)
        self.attn = _QKV(dim, heads, hd, dtype, device)
        self.attn_context = _QKV(dim, heads, hd, dtype, device)
        self.to_out = GroupQueryAttention_O(dim, hd, heads, heads, None, dtype, True, None, device)
        self.to_add_out = GroupQueryAttention_O(dim, hd, heads, heads, None, dtype, True, None, device)
        self.ff = _FF(dim, dtype, device)
        self.ff_context = _FF(dim, dtype, device)
        self.dim, self.scale = dim, hd ** -0.5

    def forward(self, x, c, temb, cos, sin, cp=None):
        xn, (g_msa, sh_mlp, sc_mlp, g_mlp) = self.norm1(x, temb)
        cn, (cg_msa, csh_mlp, csc_mlp, cg_mlp) = self.norm1_context(c, temb)
        q, k, v = self.attn(xn)
        cq, ck, cv = self.attn_context(cn)
        Nc = c.shape[1]
        q, k, v = torch.cat([cq, q], 1), torch.cat([ck, k], 1), torch.cat([cv, v], 1)
        o = _attend(_rope(q, cos, sin), _rope(k, cos, sin), v, self.scale, cp)
        B, N, H, D = o.shape
        o = o.reshape(B, N, H * D)
        x = x + g_msa * self.to_out(o[:, Nc:])
        c = c + cg_msa * self.to_add_out(o[:, :Nc])
        ln = lambda t: nn.functional.layer_norm(t, (self.dim,), eps=1e-6)  # noqa: E731
        x = x + g_mlp * self.ff(ln(x) * (1 + sc_mlp) + sh_mlp)
        c = c + cg_mlp * self.ff_context(ln(c) * (1 + csc_mlp) + csh_mlp)
        return x, c


class FluxSingleBlock(nn.Module):
    def __init__(self, dim, heads, hd, dtype, device):
        super().__init__()
        self.norm = AdaLayerNormZero(dim, 3, dtype, device)
        self.attn = _QKV(dim, heads, hd, dtype, device)
        self.proj_mlp = ColumnParallelLinear(dim, 4 * dim, bias=True, gather_output=False, dtype=dtype, device=device)
        # proj_out consumes [attention heads | mlp columns]: two row-parallel halves sharing one bias
        self.proj_out_attn = GroupQueryAttention_O(dim, hd, heads, heads, None, dtype, True, None, device)
        self.proj_out_mlp = RowParallelLinear(4 * dim, dim, bias=False, input_is_parallel=True, dtype=dtype, device=device)
        self.scale = hd ** -0.5

    def forward(self, x, temb, cos, sin, cp=None):
        xn, (gate,) = self.norm(x, temb)
        q, k, v = self.attn(xn)
        o = _attend(_rope(q, cos, sin), _rope(k, cos, sin), v, self.scale, cp)
        B, N, H, D = o.shape
        m = nn.functional.gelu(self.proj_mlp(xn), approximate="tanh")
        return x + gate * (self.proj_out_attn(o.reshape(B, N, H * D)) + self.proj_out_mlp(m))


class NeuronFluxTransformer2DModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        c, dt = config, config.neuron_config.torch_dtype
        self.config = config
        hd, heads = c.attention_head_dim, c.num_attention_heads
        dim = hd * heads
        self.inner_dim = dim
        self.x_embedder = nn.Linear(c.in_channels, dim, dtype=dt, device=device)
        self.context_embedder = nn.Linear(c.joint_attention_dim, dim, dtype=dt, device=device)
        self.time_text_embed = CombinedTimestepGuidanceTextProjEmbeddings(dim, c.pooled_projection_dim, c.guidance_embeds, dt, device)
        self.transformer_blocks = nn.ModuleList([FluxDoubleBlock(dim, heads, hd, dt, device) for _ in range(c.num_layers)])
        self.single_transformer_blocks = nn.ModuleList([FluxSingleBlock(dim, heads, hd, dt, device) for _ in range(c.num_single_layers)])
        self.norm_out = AdaLayerNormContinuous(dim, dim, dt, device)
        self.proj_out = nn.Linear(dim, c.patch_size ** 2 * c.out_channels, dtype=dt, device=device)
        self.cp_group = None            # set by the application when ``context_parallel_enabled`` (2 replicas of the TP group)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance=None):
        """hidden_states [B,N_img,C_in] packed latents; encoder_hidden_states [B,N_txt,4096] (T5); pooled [B,768] (CLIP);
        timestep / guidance [B] in [0,1] (scaled by 1000 here, as diffusers does)."""
        dt = self.x_embedder.weight.dtype
        x = self.x_embedder(hidden_states.to(dt))
        temb = self.time_text_embed(timestep.to(dt) * 1000, None if guidance is None else guidance.to(dt) * 1000,
                                    pooled_projections.to(dt))
        c = self.context_embedder(encoder_hidden_states.to(dt))
        cp, g = None, self.cp_group
        if g is not None and g.size > 1:
            n = x.shape[1] // g.size
            assert n * g.size == x.shape[1], "image tokens must divide evenly over the context-parallel ranks"
            x, img_ids = x[:, g.rank * n:(g.rank + 1) * n], img_ids[g.rank * n:(g.rank + 1) * n]
            cp = (g, c.shape[1])
        ids = torch.cat([txt_ids, img_ids], 0)
        cos, sin = rope_nd(ids, self.config.axes_dims_rope)
        cos, sin = cos.unsqueeze(0).expand(x.shape[0], -1, -1), sin.unsqueeze(0).expand(x.shape[0], -1, -1)
        for blk in self.transformer_blocks:
            x, c = blk(x, c, temb, cos, sin, cp)
        h = torch.cat([c, x], 1)
        for blk in self.single_transformer_blocks:
            h = blk(h, temb, cos, sin, cp)
        x = self.proj_out(self.norm_out(h[:, c.shape[1]:], temb))
        return x if cp is None else mappings.all_gather(x.contiguous(), 1, g)


def convert_diffusers_flux_state_dict(sd: dict, config) -> dict:
    """diffusers ``FluxTransformer2DModel`` names -> this module (q/k/v fused, proj_out of single blocks split)."""
    out = {}
    dim = config.attention_head_dim * config.num_attention_heads
    sd = dict(sd)

    def fuse(prefix, names, dst):
        for suf in ("weight", "bias"):
            out[f"{dst}.proj.{suf}"] = torch.cat([sd.pop(f"{prefix}.{n}.{suf}") for n in names], 0)
    for i in range(config.num_layers):
        p = f"transformer_blocks.{i}"
        fuse(f"{p}.attn", ("to_q", "to_k", "to_v"), f"{p}.attn")
        fuse(f"{p}.attn", ("add_q_proj", "add_k_proj", "add_v_proj"), f"{p}.attn_context")
        for a, b in (("norm_q", "attn.norm_q"), ("norm_k", "attn.norm_k"), ("norm_added_q", "attn_context.norm_q"), ("norm_added_k", "attn_context.norm_k")):
            out[f"{p}.{b}.weight"] = sd.pop(f"{p}.attn.{a}.weight")
        for suf in ("weight", "bias"):
            out[f"{p}.to_out.{suf}"] = sd.pop(f"{p}.attn.to_out.0.{suf}")
            out[f"{p}.to_add_out.{suf}"] = sd.pop(f"{p}.attn.to_add_out.{suf}")
            for ff in ("ff", "ff_context"):
                out[f"{p}.{ff}.fc1.{suf}"] = sd.pop(f"{p}.{ff}.net.0.proj.{suf}")
                out[f"{p}.{ff}.fc2.{suf}"] = sd.pop(f"{p}.{ff}.net.2.{suf}")
    for i in range(config.num_single_layers):
        p = f"single_transformer_blocks.{i}"
        fuse(f"{p}.attn", ("to_q", "to_k", "to_v"), f"{p}.attn")
        out[f"{p}.attn.norm_q.weight"] = sd.pop(f"{p}.attn.norm_q.weight")
        out[f"{p}.attn.norm_k.weight"] = sd.pop(f"{p}.attn.norm_k.weight")
        w = sd.pop(f"{p}.proj_out.weight")
        out[f"{p}.proj_out_attn.weight"], out[f"{p}.proj_out_mlp.weight"] = w[:, :dim].contiguous(), w[:, dim:].contiguous()
        out[f"{p}.proj_out_attn.bias"] = sd.pop(f"{p}.proj_out.bias")
    out.update(sd)
    return out
