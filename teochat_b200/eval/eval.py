"""Drop-in for ``videollava/eval/eval.py::load_model`` (:15-34): returns
``(tokenizer, model, processor)`` with the reference's signature.

Offline there is no checkpoint, tokenizer or network, so ``model_path`` selects
a synthetic ("random-init") model: ``"teochat-synthetic"`` / ``"teochat-synthetic-tiny"``,
optionally ``?seed=N``.  Like the reference's builder (builder.py:33) the model name must contain
``llava`` or ``teochat``.  A directory path loads an HF-format checkpoint (merged, or a LoRA adapter over ``model_base``; a
separate image-tower checkpoint directory may be given through ``cache_dir``) via teochat_b200.checkpoint.
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
    dev = torch.device(device if device is not None else "cuda:0")
    if os.path.isdir(path):
        # HF-format directory: merged checkpoint, or LoRA adapter over `model_base` (builder.py:33-112)
        from .. import checkpoint as CK
        cfg = CK.read_config(path)
        sd = CK.load_state_dict(path, model_base if (model_base and os.path.isdir(str(model_base))) else None, tower_path=cache_dir)
        weights = TeoWeights.from_state_dict(sd, cfg, dev)
        del sd
        model = TeoModel(cfg, weights, dev)
        model.model.video_tower = None
        return CK.load_tokenizer(model_base if (model_base and os.path.isdir(str(model_base))) else path, cfg.llama.vocab_size), model, \
            TeoImageProcessor(cfg.vision.image_size)
    seed = 1234
    for kv in filter(None, query.split("&")):
        k, _, v = kv.partition("=")
        if k == "seed":
            seed = int(v)
    cfg = TeoConfig.tiny() if model_name.endswith("-tiny") else TeoConfig.full()
    weights = TeoWeights.from_synthetic(cfg, seed, dev)
    model = TeoModel(cfg, weights, dev)
    model.model.video_tower = None                     # eval.py:31
    tokenizer = StubTokenizer(cfg.llama.vocab_size)
    processor = TeoImageProcessor(cfg.vision.image_size)   # processor['image'] (eval.py:33)
    return tokenizer, model, processor
