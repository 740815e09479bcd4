"""Image-size arithmetic of the Qwen2-VL processor (reference models/qwen2_vl/utils/vision_utils.py:11-70): how many patches / pixels an
image of a given size becomes after the processor's ``smart_resize`` (sides rounded to multiples of patch x merge, area clamped)."""
from __future__ import annotations

import math

from .constants import DEFAULT_IMAGE_HEIGHT, DEFAULT_IMAGE_WIDTH, MAX_PIXELS, MERGE_SIZE, MIN_PIXELS, PATCH_SIZE


def smart_resize(height: int, width: int, factor: int = PATCH_SIZE * MERGE_SIZE, min_pixels: int = MIN_PIXELS, max_pixels: int = MAX_PIXELS):
    h = max(factor, round(height / factor) * factor)
    w = max(factor, round(width / factor) * factor)
    if h * w > max_pixels:
        beta = math.sqrt(height * width / max_pixels)
        h, w = math.floor(height / beta / factor) * factor, math.floor(width / beta / factor) * factor
    elif h * w < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h, w = math.ceil(height * beta / factor) * factor, math.ceil(width * beta / factor) * factor
    return h, w


def get_image_dimensions(neuron_config):
    return (getattr(neuron_config, "default_image_width", DEFAULT_IMAGE_WIDTH), getattr(neuron_config, "default_image_height", DEFAULT_IMAGE_HEIGHT))


def calculate_max_grid_size(image_width: int, image_height: int, patch_size: int = PATCH_SIZE) -> int:
    h, w = smart_resize(image_height, image_width, factor=patch_size * MERGE_SIZE)
    return max(h // patch_size, w // patch_size)


def calculate_pixels_per_image(image_width: int, image_height: int, patch_size: int = PATCH_SIZE) -> int:
    """Rows of ``pixel_values`` one image contributes (patches after the resize)."""
    h, w = smart_resize(image_height, image_width, factor=patch_size * MERGE_SIZE)
    return (h // patch_size) * (w // patch_size)
