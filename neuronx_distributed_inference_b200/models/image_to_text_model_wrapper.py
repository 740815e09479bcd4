"""reference import path ``models.image_to_text_model_wrapper`` (:1-309): ``ImageToTextModelWrapper`` adds the vision-embedding /
vision-mask inputs (padded to the bucket, scattered at the placeholder positions) to the text sub-model wrapper.  Here the text
runner takes those as keyword tensors and pads them with the prompt (``runtime/runner.py``: "vision kw padding"), the vision tower
runs through ``EncoderRunner``."""
from ..runtime.runner import SubModelRunner as ImageToTextModelWrapper  # noqa: F401
from .encoder_base import EncoderRunner as VisionModelWrapper  # noqa: F401
from .image_to_text_model_base import ImageToTextInferenceConfig, NeuronBaseForImageToText  # noqa: F401
