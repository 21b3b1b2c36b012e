from teochat_b200.eval.metrics import detection_metrics, evaluate_masks  # noqa: F401
