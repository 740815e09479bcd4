from .hidden_state import HiddenStateRollingBuffer  # noqa: F401
from .token_tree import TokenTree  # noqa: F401
from .dynamic_token_tree import DynamicTokenTree  # noqa: F401
