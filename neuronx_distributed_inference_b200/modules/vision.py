"""Vision-transformer building blocks shared by the image encoders (Qwen2/3-VL, Pixtral, Llama-4 vision, Mllama, CLIP).

The reference builds each vision tower from its own copies of attention/MLP classes (e.g. models/qwen2_vl/modeling_qwen2_vl_vision.py,
models/pixtral/modeling_pixtral_vision.py, models/llama4/modeling_llama4_vision.py:1-1214); here one tensor-parallel block
covers them: fused QKV sharded by head, bidirectional attention restricted to the tokens of the same image (block-diagonal
``segment_ids``), optional 2-D rotary embedding, GELU / SwiGLU MLP."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from ..parallel.layers import ColumnParallelLinear, RowParallelLinear
from ..parallel.state import get_tensor_model_parallel_group
from .gqa import GroupQueryAttention_O, GroupQueryAttention_QKV

ACT = {"gelu": lambda x: nn.functional.gelu(x), "gelu_pytorch_tanh": lambda x: nn.functional.gelu(x, approximate="tanh"),
       "gelu_new": lambda x: nn.functional.gelu(x, approximate="tanh"), "gelu_fast": lambda x: nn.functional.gelu(x, approximate="tanh"),
       "quick_gelu": lambda x: x * torch.sigmoid(1.702 * x), "silu": nn.functional.silu, "relu": nn.functional.relu}


def rotate_half_apply(x, cos, sin):
    """x [..., N, H, D]; cos/sin [..., N, D] (full width, already duplicated over the two halves)."""
    d = x.shape[-1] // 2
    x1, x2 = x[..., :d], x[..., d:]
    rot = torch.cat([-x2, x1], -1)
    return (x.float() * cos.unsqueeze(-2) + rot.float() * sin.unsqueeze(-2)).to(x.dtype)


class VisionAttention(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int, bias: bool = True, dtype=torch.float32, device=None, tp_group=None,
                 o_bias: Optional[bool] = None, head_dim: Optional[int] = None):
        super().__init__()
        g = tp_group or get_tensor_model_parallel_group()
        self.head_dim = head_dim or embed_dim // num_heads
        self.qkv_proj = GroupQueryAttention_QKV(embed_dim, self.head_dim, num_heads, num_heads, g, dtype, bias, None, device)
        self.o_proj = GroupQueryAttention_O(embed_dim, self.head_dim, num_heads, num_heads, g, dtype,
                                            bias if o_bias is None else o_bias, None, device)
        self.n_heads = self.qkv_proj.n_q
        self.scale = self.head_dim ** -0.5

    def forward(self, x: torch.Tensor, cos=None, sin=None, segment_ids: Optional[torch.Tensor] = None,
                key_valid: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [B,N,C].  ``segment_ids`` [B,N]: tokens attend only within their own segment (several images packed in one
        row).  ``key_valid`` [B,N]: padding keys are hidden."""
        B, N, _ = x.shape
        H, D = self.n_heads, self.head_dim
        q, k, v = self.qkv_proj(x).view(B, N, 3 * H, D).split(H, 2)
        if cos is not None:
            if cos.shape[-1] == D // 2:       # interleaved (complex) rotary: Llama-4 vision
                q, k = ops.apply_rope(q, cos, sin, True), ops.apply_rope(k, cos, sin, True)
            else:
                q, k = rotate_half_apply(q, cos, sin), rotate_half_apply(k, cos, sin)
        if segment_ids is not None:
            mask = (segment_ids.unsqueeze(-1) == segment_ids.unsqueeze(-2)).unsqueeze(1)
        if key_valid is not None:
            kv = key_valid.bool().view(B, 1, 1, N)
            mask = kv if mask is None else mask & kv
        if mask is None:
            # one image / utterance per row, no padding: the bidirectional mode of the flash kernel (D in {64,128} on CUDA)
            o = ops.attention_prefill(q.contiguous(), k.contiguous(), v.contiguous(), self.scale, causal=False)
            return self.o_proj(o.reshape(B, N, H * D))
        o = ops.ref.attention_with_mask(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), mask, self.scale)
        return self.o_proj(o.transpose(1, 2).reshape(B, N, H * D))


class VisionMLP(nn.Module):
    """fc1 -> act -> fc2, or SwiGLU (gate/up fused) when ``gated``."""

    def __init__(self, dim: int, hidden: int, act: str = "gelu", bias: bool = True, gated: bool = False, dtype=torch.float32,
                 device=None, tp_group=None, out_dim: Optional[int] = None):
        super().__init__()
        g = tp_group or get_tensor_model_parallel_group()
        self.gated, self.act = gated, act
        if gated:
            self.gate_up_proj = ColumnParallelLinear(dim, 2 * hidden, bias=bias, gather_output=False, dtype=dtype, device=device,
                                                     tensor_model_parallel_group=g, stride=2)
        else:
            self.fc1 = ColumnParallelLinear(dim, hidden, bias=bias, gather_output=False, dtype=dtype, device=device,
                                            tensor_model_parallel_group=g)
        self.fc2 = RowParallelLinear(hidden, out_dim or dim, bias=bias, input_is_parallel=True, dtype=dtype, device=device,
                                     tensor_model_parallel_group=g)

    def forward(self, x):
        if self.gated:
            gu = self.gate_up_proj(x)
            h = gu.shape[-1] // 2
            return self.fc2(ACT[self.act](gu[..., :h]) * gu[..., h:])
        return self.fc2(ACT[self.act](self.fc1(x)))


class PatchEmbed(nn.Module):
    """Convolutional patchify expressed as a GEMM over flattened patches (the conv has stride == kernel)."""

    def __init__(self, patch_dim: int, embed_dim: int, bias: bool = False, dtype=torch.float32, device=None):
        super().__init__()
        self.proj = nn.Linear(patch_dim, embed_dim, bias=bias, dtype=dtype, device=device)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, patches):
        return self.proj(patches.to(self.proj.weight.dtype))
