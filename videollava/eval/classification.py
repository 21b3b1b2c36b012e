from teochat_b200.eval.metrics import classification_metrics  # noqa: F401
from teochat_b200.eval.metrics import normalise as _normalise


def get_string_cleaner(ignore_casing, ignore_punctuation):
    """videollava/eval/classification.py:5-12"""
    return lambda s: _normalise(s, ignore_casing, ignore_punctuation)
