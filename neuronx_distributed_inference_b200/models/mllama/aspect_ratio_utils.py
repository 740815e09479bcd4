"""Import path of the reference (models/mllama/aspect_ratio_utils.py)."""
from .utils import convert_aspect_ratios_to_ids, get_all_supported_aspect_ratios, get_aspect_ratio_mask  # noqa: F401
