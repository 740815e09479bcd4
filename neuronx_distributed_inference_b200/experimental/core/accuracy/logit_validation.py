"""reference experimental/core/accuracy/logit_validation.py — the production implementation lives in :mod:`....utils.accuracy`."""
from ....utils.accuracy import check_accuracy_logits, logit_validation  # noqa: F401
