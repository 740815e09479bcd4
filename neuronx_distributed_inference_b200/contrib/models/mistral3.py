"""Mistral-Small-3.1 (``Mistral3ForConditionalGeneration``): the Pixtral vision tower and Mistral text decoder of
``models/pixtral`` with the Mistral-3 projector in between — RMSNorm, a learned ``spatial_merge_size x spatial_merge_size`` patch
merger (unfold + linear) that cuts the image tokens by 4, then the two-layer GELU MLP.
reference port: contrib/models/Mistral-Small-3.1-24B-Instruct-2503/src/modeling_mistral3.py."""
from __future__ import annotations

import torch
import torch.nn as nn

from ...models.pixtral.modeling_pixtral import ACT, NeuronPixtralForCausalLM, NeuronPixtralVisionModel
from ...modules.norm import RMSNorm


class NeuronMistral3VisionModel(NeuronPixtralVisionModel):
    def __init__(self, config, device=None):
        super().__init__(config, device)
        vc, tc = config.vision_config, config.get_text_config()
        dt = vc.neuron_config.torch_dtype
        self.merge = int(getattr(config, "spatial_merge_size", 2))
        nfeat = 1 if isinstance(self.feature_layer, int) else len(self.feature_layer)
        bias = getattr(config, "multimodal_projector_bias", False)
        self.proj_norm = RMSNorm(vc.hidden_size, getattr(tc, "rms_norm_eps", 1e-5), dt, device=device)
        self.merging_layer = nn.Linear(vc.hidden_size * self.merge ** 2, vc.hidden_size, bias=False, dtype=dt, device=device)
        self.proj1 = nn.Linear(vc.hidden_size * nfeat, tc.hidden_size, bias=bias, dtype=dt, device=device)
        self.proj2 = nn.Linear(tc.hidden_size, tc.hidden_size, bias=bias, dtype=dt, device=device)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, pixel_values, image_sizes=None):
        feat, grids = self.features(pixel_values, image_sizes)
        feat = self.proj_norm(feat)
        s, d, out, off = self.merge, feat.shape[-1], [], 0
        for gh, gw in grids:
            g = feat[off:off + gh * gw].view(gh, gw, d)
            off += gh * gw
            # s x s neighbourhoods, channel-major inside a neighbourhood (the layout torch.nn.functional.unfold produces)
            g = g.view(gh // s, s, gw // s, s, d).permute(0, 2, 4, 1, 3).reshape((gh // s) * (gw // s), d * s * s)
            out.append(g)
        x = self.merging_layer(torch.cat(out, 0))
        return self.proj2(ACT[self.proj_act](self.proj1(x)))


class NeuronMistral3ForCausalLM(NeuronPixtralForCausalLM):
    _vision_cls = NeuronMistral3VisionModel

    def _split_state_dict(self, sd):
        ren = (("multi_modal_projector.norm.", "vision_tower.proj_norm."),
               ("multi_modal_projector.patch_merger.merging_layer.", "vision_tower.merging_layer."))
        out = {}
        for k, v in sd.items():
            for a, b in ren:
                if k.startswith(a):
                    k = b + k[len(a):]
            out[k] = v
        return super()._split_state_dict(out)
