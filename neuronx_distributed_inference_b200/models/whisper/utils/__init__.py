from .decoding import DecodingOptions, DecodingResult, decode  # noqa: F401
