"""Qwen2.5-VL (reference contrib/models/Qwen2.5-VL-{3B,32B}-Instruct): Qwen2-VL text stack (M-RoPE) with a different vision
tower — RMSNorm, biased SwiGLU MLP, **window attention** (all but the ``fullatt_block_indexes`` layers attend inside
``window_size``-pixel windows; tokens are re-ordered window-major once and the order is undone after the merger)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...models.qwen2_vl.modeling_qwen2_vl import NeuronQwen2VLForCausalLM, NeuronQwen2VLTextModel, Qwen2VLInferenceConfig
from ...modules.norm import RMSNorm
from ...modules.vision import PatchEmbed, VisionAttention, VisionMLP


class Qwen25VLBlock(nn.Module):
    def __init__(self, vc, dtype, device):
        super().__init__()
        self.norm1 = RMSNorm(vc.hidden_size, 1e-6, dtype, device=device)
        self.norm2 = RMSNorm(vc.hidden_size, 1e-6, dtype, device=device)
        self.attn = VisionAttention(vc.hidden_size, vc.num_heads, True, dtype, device)
        self.mlp = VisionMLP(vc.hidden_size, vc.intermediate_size, getattr(vc, "hidden_act", "silu"), True, True, dtype, device)

    def forward(self, x, cos, sin, seg):
        x = x + self.attn(self.norm1(x), cos, sin, seg)
        return x + self.mlp(self.norm2(x))


class NeuronQwen25VLVisionModel(nn.Module):
    def __init__(self, config, device=None):
        super().__init__()
        vc = config.vision_config
        dt = vc.neuron_config.torch_dtype
        self.vc, self.merge = vc, vc.spatial_merge_size
        self.unit = self.merge * self.merge
        self.patch_embed = PatchEmbed(vc.in_channels * vc.temporal_patch_size * vc.patch_size ** 2, vc.hidden_size, False, dt, device)
        self.blocks = nn.ModuleList([Qwen25VLBlock(vc, dt, device) for _ in range(vc.depth)])
        self.full_layers = set(getattr(vc, "fullatt_block_indexes", []))
        self.ln_q = RMSNorm(vc.hidden_size, 1e-6, dt, device=device)
        hid = vc.hidden_size * self.unit
        self.merger_fc1 = nn.Linear(hid, hid, dtype=dt, device=device)
        self.merger_fc2 = nn.Linear(hid, vc.out_hidden_size, dtype=dt, device=device)
        self.head_dim = vc.hidden_size // vc.num_heads
        for p in self.parameters():
            p.requires_grad_(False)

    def _layout(self, grid_thw, device):
        """-> rotary (cos, sin) per patch, window permutation over merge units, per-unit image ids and window ids (window order)."""
        m, vc = self.merge, self.vc
        win = vc.window_size // m // vc.patch_size
        ids, order, img_of_unit, win_of_unit = [], [], [], []
        base = wid = 0
        for img, (t, h, w) in enumerate(grid_thw.tolist()):
            hp = torch.arange(h).view(h, 1).expand(h, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
            wp = torch.arange(w).view(1, w).expand(h, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
            ids.append(torch.stack([hp, wp], -1).repeat(t, 1))
            gh, gw = h // m, w // m
            idx = torch.arange(t * gh * gw).view(t, gh, gw)
            ph, pw = win - gh % win, win - gw % win
            nh, nw = (gh + ph) // win, (gw + pw) // win
            pad = torch.nn.functional.pad(idx, (0, pw, 0, ph), value=-100).view(t, nh, win, nw, win).permute(0, 1, 3, 2, 4).reshape(t, nh * nw, win * win)
            for f in range(t):
                for k in range(nh * nw):
                    sel = pad[f, k][pad[f, k] != -100]
                    if sel.numel():
                        order.append(sel + base)
                        win_of_unit.append(torch.full((sel.numel(),), wid))
                        img_of_unit.append(torch.full((sel.numel(),), img * 100000 + f))
                        wid += 1
            base += t * gh * gw
        ids = torch.cat(ids).to(device)
        dim = self.head_dim // 2
        inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32, device=device) / dim))
        fr = torch.outer(torch.arange(int(grid_thw[:, 1:].max()), dtype=torch.float32, device=device), inv)
        emb = fr[ids].flatten(1)
        emb = torch.cat([emb, emb], -1)
        return emb, torch.cat(order).to(device), torch.cat(img_of_unit).to(device), torch.cat(win_of_unit).to(device)

    def forward(self, pixel_values, image_grid_thw):
        x = self.patch_embed(pixel_values)
        N, C = x.shape
        emb, order, img_u, win_u = self._layout(image_grid_thw.cpu(), x.device)
        U = self.unit
        x = x.view(N // U, U, C)[order].reshape(N, C)
        emb = emb.view(N // U, U, -1)[order].reshape(N, -1)
        seg_full = img_u.repeat_interleave(U).to(torch.int32).unsqueeze(0)
        seg_win = win_u.repeat_interleave(U).to(torch.int32).unsqueeze(0)
        cos, sin = emb.cos().unsqueeze(0), emb.sin().unsqueeze(0)
        x = x.unsqueeze(0)
        for i, blk in enumerate(self.blocks):
            x = blk(x, cos, sin, seg_full if i in self.full_layers else seg_win)
        y = self.ln_q(x.squeeze(0)).view(N // U, U * C)
        y = self.merger_fc2(nn.functional.gelu(self.merger_fc1(y)))
        return y[torch.argsort(order)]


class NeuronQwen25VLForCausalLM(NeuronQwen2VLForCausalLM):
    _model_cls = NeuronQwen2VLTextModel
    _vision_cls = NeuronQwen25VLVisionModel

    @classmethod
    def get_config_cls(cls):
        return Qwen2VLInferenceConfig

    @staticmethod
    def convert_hf_to_neuron_vision_state_dict(sd, config):
        out = {}
        for k, v in sd.items():
            if k == "patch_embed.proj.weight":
                v = v.reshape(v.shape[0], -1)
            k = (k.replace(".attn.qkv.", ".attn.qkv_proj.").replace(".attn.proj.", ".attn.o_proj.").replace(".mlp.down_proj.", ".mlp.fc2.")
                 .replace("merger.ln_q.", "ln_q.").replace("merger.mlp.0.", "merger_fc1.").replace("merger.mlp.2.", "merger_fc2."))
            out[k] = v
        for i in range(config.vision_config.depth):
            for suf in ("weight", "bias"):
                g, u = f"blocks.{i}.mlp.gate_proj.{suf}", f"blocks.{i}.mlp.up_proj.{suf}"
                if g in out:
                    out[f"blocks.{i}.mlp.gate_up_proj.{suf}"] = torch.cat([out.pop(g), out.pop(u)], 0)
        return out
