from .registry import TensorReplacementRegistry, replace_tensors  # noqa: F401
