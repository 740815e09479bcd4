"""Variable-size image tiling for Mllama (reference models/mllama/image_transform.py:27-470 ``VariableSizeImageTransform``): pick the
tile arrangement (canvas) that fits the image best, resize without distortion, pad to the canvas, cut into ``tile x tile`` crops.
Tensor-only (no PIL dependency): images are ``[3, H, W]`` float tensors in [0, 1]."""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.nn.functional as F

from .utils import get_all_supported_aspect_ratios


class VariableSizeImageTransform:
    def __init__(self, size: int = 560, max_num_tiles: int = 4, image_mean=(0.48145466, 0.4578275, 0.40821073),
                 image_std=(0.26862954, 0.26130258, 0.27577711)):
        self.size, self.max_num_tiles = size, max_num_tiles
        self.mean = torch.tensor(image_mean).view(3, 1, 1)
        self.std = torch.tensor(image_std).view(3, 1, 1)

    # ---- canvas choice ---------------------------------------------------------------------------------------------------------
    def get_optimal_tiled_canvas(self, image_height: int, image_width: int) -> Tuple[int, int]:
        """Among all canvases of at most ``max_num_tiles`` tiles: prefer the smallest UP-scaling factor (>= 1); if every canvas would
        shrink the image, the largest down-scaling factor; ties go to the canvas with the smaller area."""
        t = self.size
        cands = [(h * t, w * t) for w, h in get_all_supported_aspect_ratios(self.max_num_tiles)]
        scales = [min(ch / image_height, cw / image_width) for ch, cw in cands]
        ups = [s for s in scales if s >= 1]
        pick = min(ups) if ups else max(scales)
        best = [c for c, s in zip(cands, scales) if s == pick]
        return min(best, key=lambda c: c[0] * c[1])

    def get_image_size_fit_to_canvas(self, image_height: int, image_width: int, canvas_height: int, canvas_width: int) -> Tuple[int, int]:
        """Aspect-preserving size inside the canvas; small images are up-scaled to at most one tile more than they need."""
        t = self.size
        target_w = min(max(image_width, t), canvas_width)
        target_h = min(max(image_height, t), canvas_height)
        sh, sw = target_h / image_height, target_w / image_width
        if sw < sh:
            return min(math.floor(image_height * sw) or 1, target_h), target_w
        return target_h, min(math.floor(image_width * sh) or 1, target_w)

    # ---- pixels ---------------------------------------------------------------------------------------------------------------
    def split_to_tiles(self, image: torch.Tensor, tiles_h: int, tiles_w: int) -> torch.Tensor:
        C, H, W = image.shape
        th, tw = H // tiles_h, W // tiles_w
        return image.view(C, tiles_h, th, tiles_w, tw).permute(1, 3, 0, 2, 4).reshape(tiles_h * tiles_w, C, th, tw)

    def __call__(self, image: torch.Tensor):
        """-> (tiles ``[n_tiles, 3, size, size]`` normalised, (tiles_h, tiles_w))."""
        C, H, W = image.shape
        ch, cw = self.get_optimal_tiled_canvas(H, W)
        nh, nw = self.get_image_size_fit_to_canvas(H, W, ch, cw)
        img = F.interpolate(image.unsqueeze(0).float(), size=(nh, nw), mode="bilinear", align_corners=False, antialias=True)[0]
        img = (img - self.mean) / self.std
        img = F.pad(img, (0, cw - nw, 0, ch - nh))
        ar = (ch // self.size, cw // self.size)
        return self.split_to_tiles(img, *ar), ar


def _stack_images(tile_lists: List[List[torch.Tensor]], max_num_tiles: int, max_images: int = None):
    """Per-row lists of per-image tile tensors -> ``[B, max_images, max_tiles, 3, S, S]`` + ``[B, max_images]`` tile counts."""
    B = len(tile_lists)
    n = max_images or max(len(r) for r in tile_lists)
    S = next(t for r in tile_lists for t in r).shape[-1]
    out = torch.zeros(B, n, max_num_tiles, 3, S, S)
    counts = torch.zeros(B, n, dtype=torch.int64)
    for b, row in enumerate(tile_lists):
        for i, t in enumerate(row):
            out[b, i, : t.shape[0]] = t
            counts[b, i] = t.shape[0]
    return out, counts


def custom_image_preprocessing(images: List[List[torch.Tensor]], size: int = 560, max_num_tiles: int = 4):
    """Rows of images -> (pixel_values, aspect_ratio_ids, aspect_ratio_mask, num_tiles) ready for ``NeuronMllamaForCausalLM``."""
    from .utils import convert_aspect_ratios_to_ids, get_aspect_ratio_mask
    tf = VariableSizeImageTransform(size, max_num_tiles)
    tiles, ars = [], []
    for row in images:
        r = [tf(img) for img in row]
        tiles.append([t for t, _ in r])
        ars.append([a for _, a in r])
    pix, counts = _stack_images(tiles, max_num_tiles)
    return pix, convert_aspect_ratios_to_ids(ars, max_num_tiles), get_aspect_ratio_mask(ars, max_num_tiles), counts
