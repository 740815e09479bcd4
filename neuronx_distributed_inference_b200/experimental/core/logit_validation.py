"""Logit validation (reference experimental/core/accuracy/logit_validation.py:73-346, defaults :14-21): teacher-forced comparison
of device logits against golden logits with a top-k tolerance map and a divergence tolerance.  The production implementation
lives in :mod:`...utils.accuracy`; this module keeps the experimental entry point."""
from ...utils.accuracy import check_accuracy_logits, logit_validation  # noqa: F401
