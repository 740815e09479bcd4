"""Padding helpers of the Llama-4 vision path (reference models/llama4/utils/encoder_utils.py): the vision encoder runs on a fixed number
of image chunks and the text model takes a fixed-width block of vision embeddings plus the positions they are scattered to."""
from __future__ import annotations

import torch


def pad_image_tensor(pixel_values: torch.Tensor, num_chunks: int):
    """[n, C, H, W] -> ([num_chunks, C, H, W] zero padded, n)."""
    n = pixel_values.shape[0]
    if n > num_chunks:
        raise ValueError(f"{n} image chunks exceed the encoder bucket of {num_chunks}")
    out = pixel_values.new_zeros((num_chunks,) + tuple(pixel_values.shape[1:]))
    out[:n] = pixel_values
    return out, n


def depad_output(x: torch.Tensor, n: int) -> torch.Tensor:
    return x[:n]


def generate_positions_from_mask(vision_mask: torch.Tensor) -> torch.Tensor:
    """Boolean mask over the flattened prompt -> indices of the image-token positions."""
    return vision_mask.reshape(-1).nonzero().flatten()


def pad_positions(positions: torch.Tensor, target: int, fill: int) -> torch.Tensor:
    out = positions.new_full((target,), fill)
    out[: positions.numel()] = positions
    return out


def pad_vision_embeddings(vision_embeddings: torch.Tensor, target: int) -> torch.Tensor:
    flat = vision_embeddings.reshape(-1, vision_embeddings.shape[-1])
    out = flat.new_zeros(target, flat.shape[-1])
    out[: flat.shape[0]] = flat
    return out


def pad_image_mask(vision_mask: torch.Tensor, target: int) -> torch.Tensor:
    out = vision_mask.new_zeros(vision_mask.shape[0], target)
    out[:, : vision_mask.shape[1]] = vision_mask
    return out


def scatter_by_index_put(h: torch.Tensor, vision_embeddings: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
    """Write ``vision_embeddings[i]`` at flat position ``positions[i]`` of ``h [B, T, H]`` (positions past B*T are padding and ignored)."""
    B, T, H = h.shape
    flat = h.reshape(B * T, H).clone()
    ok = positions < B * T
    flat[positions[ok]] = vision_embeddings.reshape(-1, H)[: positions.numel()][ok].to(flat.dtype)
    return flat.view(B, T, H)


def generate_llama4_vision_encoder_buckets(dp_degree: int, max_chunks: int):
    """Powers of two up to ``max_chunks`` image chunks, each a multiple of the encoder's data-parallel degree."""
    out, b = [], max(1, dp_degree)
    while b < max_chunks:
        out.append(b)
        b *= 2
    return out + [max_chunks]
