"""reference import path ``models.model_wrapper``: ``ModelWrapper`` (per-sub-model bucketing / padding / execution) is
``runtime.runner.SubModelRunner`` here; the sub-model tags are unchanged."""
from ..runtime.runner import SubModelRunner as ModelWrapper  # noqa: F401
from .application_base import (CONTEXT_ENCODING_MODEL_TAG, FUSED_SPECULATION_MODEL_TAG, MEDUSA_MODEL_TAG, SPECULATION_MODEL_TAG,  # noqa: F401
                               TOKEN_GENERATION_MODEL_TAG)
from .encoder_base import VISION_ENCODER_MODEL_TAG  # noqa: F401
