"""Engine exceptions (reference utils/exceptions.py)."""


class LogitMatchingValidationError(AssertionError):
    def __init__(self, message="logit validation failed", results=None):
        super().__init__(message)
        self.results = results


class KernelUnavailableError(RuntimeError):
    """A CUDA op was requested on a GPU but the compiled extension is missing (never silently fall back on a GPU box)."""


class UnsupportedConfigError(ValueError):
    pass
