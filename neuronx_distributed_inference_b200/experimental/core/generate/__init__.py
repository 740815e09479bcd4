from .generate import GenerateResult, generate  # noqa: F401
