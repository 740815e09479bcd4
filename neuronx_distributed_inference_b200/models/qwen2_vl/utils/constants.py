"""Defaults of the Qwen2-VL processors (reference models/qwen2_vl/utils/constants.py)."""
DEFAULT_IMAGE_WIDTH = 640
DEFAULT_IMAGE_HEIGHT = 320
PATCH_SIZE = 14
MERGE_SIZE = 2
MIN_PIXELS = 56 * 56
MAX_PIXELS = 14 * 14 * 4 * 1280
