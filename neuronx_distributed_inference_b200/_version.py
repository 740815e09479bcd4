"""Package version (reference ``_version.py``); ``utils/version_utils.py`` reports it next to the torch / CUDA versions."""
__version__ = "0.1.0"
