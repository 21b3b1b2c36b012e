"""Drop-in for ``videollava/eval/eval.py::load_model`` (:15-34): returns
``(tokenizer, model, processor)`` with the reference's signature.

Offline there is no checkpoint, tokenizer or network, so ``model_path`` selects
a synthetic ("random-init") model: ``"teochat-synthetic"`` / ``"teochat-synthetic-tiny"``,
optionally ``?seed=N``.  Like the reference's builder (builder.py:33) the model name must contain
``llava`` or ``teochat``.  Loading a real HF checkpoint directory is SURVEY.md §8(f) row 1.
"""
from __future__ import annotations

import os

import torch

from ..config import TeoConfig
from ..engine import TeoModel
from ..mm_utils import get_model_name_from_path
from ..processor import TeoImageProcessor
from ..tokenizer import StubTokenizer
from ..weights import TeoWeights


def load_model(model_path, model_base=None, load_8bit=False, load_4bit=False, cache_dir=None, device=None):
    if load_8bit or load_4bit:
        raise NotImplementedError("bitsandbytes int8/int4 loading is out of scope (SURVEY.md §8a quirk 9); weights are bf16")
    path, _, query = str(model_path).partition("?")
    model_name = get_model_name_from_path(path)
    if "llava" not in model_name.lower() and "teochat" not in model_name.lower():
        raise ValueError(f"model name {model_name!r} must contain 'llava' or 'teochat' (builder.py:33)")
    if os.path.isdir(path):
        raise NotImplementedError("real-checkpoint loading is SURVEY.md §8(f) row 1 (next); use 'teochat-synthetic'")
    seed = 1234
    for kv in filter(None, query.split("&")):
        k, _, v = kv.partition("=")
        if k == "seed":
            seed = int(v)
    cfg = TeoConfig.tiny() if model_name.endswith("-tiny") else TeoConfig.full()
    dev = torch.device(device if device is not None else "cuda:0")
    weights = TeoWeights.from_synthetic(cfg, seed, dev)
    model = TeoModel(cfg, weights, dev)
    model.model.video_tower = None                     # eval.py:31
    tokenizer = StubTokenizer(cfg.llama.vocab_size)
    processor = TeoImageProcessor(cfg.vision.image_size)   # processor['image'] (eval.py:33)
    return tokenizer, model, processor
