"""reference models/gpt_oss/hf_configuration.py carried its own ``GptOssConfig`` while transformers did not have one; it does now."""
from transformers import GptOssConfig  # noqa: F401
