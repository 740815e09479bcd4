from .bucketing_on_seq_len import build_for_bucketing_on_seq_len  # noqa: F401
