from teochat_b200.eval.eval import load_model  # noqa: F401
