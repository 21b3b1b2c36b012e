"""Import shim: ``from videollava.eval.eval import load_model`` and
``from videollava.eval.inference import run_inference_single`` (README.md:112-125 of the
reference) resolve to the B200-native implementation in ``teochat_b200``."""
