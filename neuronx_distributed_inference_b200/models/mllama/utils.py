"""Host-side helpers of the Mllama pipeline (reference models/mllama/utils.py:23-130): which text tokens each image is visible to,
the dense cross-attention mask built from those ranges, aspect-ratio bookkeeping of the tiled vision encoder, prompt decoration."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def get_negative_inf_value(dtype) -> float:
    return torch.finfo(dtype).min


def get_input_shape(input_ids: torch.Tensor) -> Tuple[int, int]:
    return (1, input_ids.shape[0]) if input_ids.dim() == 1 else tuple(input_ids.shape[:2])


def create_vision_mask(input_ids: Sequence[int], image_token_id: int) -> List[List[int]]:
    """``[start, end)`` text ranges, one per image token: an image is visible from its token up to the next image token; runs of
    consecutive image tokens share the range of the last one; the final image sees the rest of the sequence (a single image: -1 =
    "to the end", the convention of the Hugging Face processor)."""
    locs = [i for i, t in enumerate(input_ids) if int(t) == image_token_id]
    if not locs:
        return []
    if len(locs) == 1:
        return [[locs[0], -1]]
    spans = [[a, b] for a, b in zip(locs[:-1], locs[1:])] + [[locs[-1], len(input_ids)]]
    end = spans[-1][1]
    for s in reversed(spans):
        if s[0] == s[1] - 1:
            s[1] = end
        end = s[1]
    return spans


def vision_mask_to_dense(spans_per_row: List[List[List[int]]], num_tiles_per_row: List[List[int]], seq_len: int, max_images: int,
                         max_tiles: int) -> torch.Tensor:
    """-> ``[B, seq_len, max_images, max_tiles]`` 0/1 cross-attention mask (the model flattens the last two axes times the patches per
    tile)."""
    B = len(spans_per_row)
    out = torch.zeros(B, seq_len, max_images, max_tiles, dtype=torch.int64)
    for b, (spans, tiles) in enumerate(zip(spans_per_row, num_tiles_per_row)):
        for i, ((s, e), n) in enumerate(zip(spans, tiles)):
            e = seq_len if e == -1 else min(e, seq_len)
            out[b, s:e, i, :n] = 1
    return out


def get_all_supported_aspect_ratios(max_image_tiles: int) -> List[Tuple[int, int]]:
    """(tiles along width, tiles along height) with at most ``max_image_tiles`` tiles, width-major order: id = index + 1."""
    return [(w, h) for w in range(1, max_image_tiles + 1) for h in range(1, max_image_tiles + 1) if w * h <= max_image_tiles]


def convert_aspect_ratios_to_ids(aspect_ratios: List[List[Tuple[int, int]]], max_image_tiles: int) -> torch.Tensor:
    """Per image ``(tiles_h, tiles_w)`` -> id (0 = padding image)."""
    table = get_all_supported_aspect_ratios(max_image_tiles)
    B, n = len(aspect_ratios), max(len(r) for r in aspect_ratios)
    ids = torch.zeros(B, n, dtype=torch.int64)
    for b, row in enumerate(aspect_ratios):
        for i, (th, tw) in enumerate(row):
            ids[b, i] = table.index((th, tw)) + 1
    return ids


def get_aspect_ratio_mask(aspect_ratios: List[List[Tuple[int, int]]], max_image_tiles: int) -> torch.Tensor:
    """``[B, n_images, max_tiles]``: 1 for the tiles an image really has; padding images keep their first tile switched on."""
    B, n = len(aspect_ratios), max(len(r) for r in aspect_ratios)
    m = torch.zeros(B, n, max_image_tiles, dtype=torch.int64)
    m[:, :, 0] = 1
    for b, row in enumerate(aspect_ratios):
        for i, (th, tw) in enumerate(row):
            m[b, i, : th * tw] = 1
    return m


def add_instruct(prompt: str, has_image: bool) -> str:
    """Llama-3.2-Vision instruct decoration of a bare user prompt."""
    img = "<|image|>" if has_image else ""
    return f"<|begin_of_text|><|start_header_id|>user<|end_header_id|>\\n\\n{img}{prompt}<|eot_id|><|start_header_id|>assistant<|end_header_id|>\\n\\n"
