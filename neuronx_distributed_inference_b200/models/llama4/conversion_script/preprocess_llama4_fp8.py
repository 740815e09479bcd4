"""Offline conversion of the Hugging Face Llama-4 FP8 checkpoints (Maverick 128E) to the layout the Llama-4 loader reads
(reference models/llama4/conversion_script/preprocess_llama4_FP8.py).

The released FP8 checkpoint stores every routed expert separately — ``feed_forward.experts.<e>.{gate,up,down}_proj.weight`` in
``float8_e4m3fn`` with a per-output-channel ``weight_scale`` — while the bf16 checkpoints (and the loader) use the fused tensors
``experts.gate_up_proj [E, H, 2I]`` / ``experts.down_proj [E, I, H]``.  This script fuses the experts layer by layer.

Differences from the reference: Trainium's FP8 is IEEE E4M3 (max 240) so the reference rescales every weight by 448/240; Blackwell's
tensor cores use the same OCP ``e4m3fn`` the checkpoint ships, so nothing is rescaled.  The routed experts are de-quantised to bf16 here
(``--keep-fp8`` writes the fused fp8 tensors with fused ``.scale`` entries instead, for the quantised-expert path)."""
from __future__ import annotations

import argparse
import json
import os
from typing import Dict

import torch


def fuse_layer_experts(sd: Dict[str, torch.Tensor], prefix: str, num_experts: int, keep_fp8: bool = False, out_dtype=torch.bfloat16) -> bool:
    """Fuse ``<prefix>feed_forward.experts.<e>.*`` in place.  Returns False when the layer has no per-expert tensors (dense layer or an
    already fused checkpoint)."""
    base = prefix + "feed_forward.experts."
    if f"{base}0.gate_proj.weight" not in sd:
        return False

    def take(e, name):
        w = sd.pop(f"{base}{e}.{name}.weight")
        s = sd.pop(f"{base}{e}.{name}.weight_scale", None)
        return w, (None if s is None else s.float().reshape(-1, 1))

    gu_w, gu_s, dn_w, dn_s = [], [], [], []
    for e in range(num_experts):
        (gw, gs), (uw, us), (dw, ds) = take(e, "gate_proj"), take(e, "up_proj"), take(e, "down_proj")
        if keep_fp8:
            gu_w.append(torch.cat([gw.view(torch.uint8), uw.view(torch.uint8)], 0).view(gw.dtype) if gw.dtype != out_dtype else torch.cat([gw, uw], 0))
            gu_s.append(torch.cat([gs, us], 0))
            dn_w.append(dw)
            dn_s.append(ds)
        else:
            dq = lambda w, s: (w.float() * (1.0 if s is None else s)).to(out_dtype)               # noqa: E731
            gu_w.append(torch.cat([dq(gw, gs), dq(uw, us)], 0))                                    # [2I, H]
            dn_w.append(dq(dw, ds))                                                                # [H, I]
    if keep_fp8:
        stack = lambda ts: torch.stack([t.view(torch.uint8) for t in ts]).view(ts[0].dtype) if ts[0].dtype.itemsize == 1 else torch.stack(ts)  # noqa: E731
        sd[base + "gate_up_proj"] = stack(gu_w).transpose(1, 2).contiguous()                        # [E, H, 2I]
        sd[base + "gate_up_proj.scale"] = torch.stack(gu_s).transpose(1, 2).contiguous()            # [E, 1, 2I]
        sd[base + "down_proj"] = stack(dn_w).transpose(1, 2).contiguous()                           # [E, I, H]
        sd[base + "down_proj.scale"] = torch.stack(dn_s).transpose(1, 2).contiguous()               # [E, 1, H]
    else:
        sd[base + "gate_up_proj"] = torch.stack(gu_w).transpose(1, 2).contiguous()
        sd[base + "down_proj"] = torch.stack(dn_w).transpose(1, 2).contiguous()
    return True


def dequantize_dense(sd: Dict[str, torch.Tensor], out_dtype=torch.bfloat16) -> int:
    """Any remaining ``<name>.weight`` / ``<name>.weight_scale`` pair (shared expert, attention) -> bf16 weight."""
    n = 0
    for k in [k for k in sd if k.endswith(".weight_scale")]:
        w = k[: -len("_scale")]
        if w in sd and sd[w].dtype.itemsize == 1:
            sd[w] = (sd[w].float() * sd.pop(k).float().reshape(-1, 1)).to(out_dtype)
            n += 1
    return n


def convert(hf_fp8_model_path: str, save_model_path: str, keep_fp8: bool = False):
    from ....modules.checkpoint import load_state_dict, save_state_dict_safetensors
    with open(os.path.join(hf_fp8_model_path, "config.json")) as f:
        cfg = json.load(f)
    tc = cfg.get("text_config", cfg)
    sd = load_state_dict(hf_fp8_model_path)
    fused = 0
    for i in range(tc["num_hidden_layers"]):
        for prefix in (f"language_model.model.layers.{i}.", f"model.layers.{i}.", f"layers.{i}."):
            fused += fuse_layer_experts(sd, prefix, tc["num_local_experts"], keep_fp8)
    dense = 0 if keep_fp8 else dequantize_dense(sd)
    os.makedirs(save_model_path, exist_ok=True)
    save_state_dict_safetensors(sd, save_model_path)
    cfg.pop("quantization_config", None) if not keep_fp8 else None
    with open(os.path.join(save_model_path, "config.json"), "w") as f:
        json.dump(cfg, f, indent=2)
    print(f"fused the experts of {fused} MoE layers, de-quantised {dense} dense projections -> {save_model_path}")


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\\n\\n")[0])
    ap.add_argument("hf_fp8_model_path")
    ap.add_argument("save_model_path")
    ap.add_argument("--keep-fp8", action="store_true")
    a = ap.parse_args()
    convert(a.hf_fp8_model_path, a.save_model_path, a.keep_fp8)


if __name__ == "__main__":
    main()
