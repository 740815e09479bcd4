"""MXFP4 (OCP microscaling: 32-element blocks of fp4 e2m1 sharing one E8M0 scale) pack / unpack for the GPT-OSS experts.

reference: models/gpt_oss/mx_layout_transform.py (767 LoC: ``pack_fp4_x4_uint16`` and the hidden/intermediate shuffles that
lay the blocks out for the Neuron engines).  On B200 the released checkpoint layout — ``*_blocks`` uint8 ``[..., K/32, 16]``
(two nibbles per byte, low nibble first) and ``*_scales`` uint8 ``[..., K/32]`` (biased by 127) — is already what
``tcgen05.mma.kind::mxf4`` block-scaled operands want (K-major, 1x32 scale granularity), so no shuffle is needed; this module
provides the (de)quantisers used at load time (bf16 expert weights) and by tests."""
from __future__ import annotations

import torch

FP4_VALUES = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0, -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0]


def dequantize_mxfp4(blocks: torch.Tensor, scales: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """blocks uint8 [..., G, 16], scales uint8 [..., G] -> [..., G*32] (``dtype``)."""
    lut = torch.tensor(FP4_VALUES, dtype=torch.float32, device=blocks.device)
    lo = lut[(blocks & 0x0F).long()]
    hi = lut[(blocks >> 4).long()]
    vals = torch.stack([lo, hi], -1).flatten(-2)                     # [..., G, 32]
    exp = scales.to(torch.int32) - 127
    out = torch.ldexp(vals, exp.unsqueeze(-1))
    return out.flatten(-2).to(dtype)


def quantize_mxfp4(w: torch.Tensor):
    """[..., K] (K % 32 == 0) -> (blocks uint8 [..., K/32, 16], scales uint8 [..., K/32]); round-to-nearest on the e2m1 grid."""
    *lead, K = w.shape
    x = w.float().reshape(*lead, K // 32, 32)
    amax = x.abs().amax(-1, keepdim=True).clamp_min(2.0 ** -126)
    exp = torch.floor(torch.log2(amax)) - 2                           # largest magnitude lands in [4, 8) -> clamps to 6
    scaled = x / torch.exp2(exp)
    grid = torch.tensor(FP4_VALUES[:8], device=w.device)
    idx = (scaled.abs().unsqueeze(-1) - grid).abs().argmin(-1)
    code = idx + (scaled < 0).long() * 8
    blocks = (code[..., 0::2] | (code[..., 1::2] << 4)).to(torch.uint8)
    scales = (exp.squeeze(-1) + 127).clamp(0, 254).to(torch.uint8)
    return blocks, scales


def pack_fp4_x4_uint16(codes: torch.Tensor) -> torch.Tensor:
    """Four fp4 codes per uint16 (element i in bits 4i..4i+3) — the packing the reference hands to its MX matmul."""
    c = codes.to(torch.int32).reshape(*codes.shape[:-1], -1, 4)
    return (c[..., 0] | (c[..., 1] << 4) | (c[..., 2] << 8) | (c[..., 3] << 12)).to(torch.int32)


def dequantize_mxfp4_state_dict(sd: dict, dtype=torch.bfloat16) -> dict:
    """Replace every ``<name>_blocks`` / ``<name>_scales`` pair by ``<name>`` in the layout transformers exposes after its
    own dequantisation (``[E, in, out]``: the packed tensors are ``[E, out, in/32, 16]``)."""
    out = dict(sd)
    for k in [k for k in sd if k.endswith("_blocks")]:
        base = k[: -len("_blocks")]
        sk = base + "_scales"
        if sk not in sd:
            continue
        w = dequantize_mxfp4(out.pop(k), out.pop(sk), dtype)          # [E, out, in]
        out[base] = w.transpose(1, 2).contiguous()
    return out
