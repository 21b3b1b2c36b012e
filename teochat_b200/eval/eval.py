"""Drop-in for ``videollava/eval/eval.py::load_model`` (:15-34): returns
``(tokenizer, model, processor)`` with the reference's signature.

Offline there is no checkpoint, tokenizer or network, so ``model_path`` selects
a synthetic ("random-init") model: ``"teochat-synthetic"`` / ``"teochat-synthetic-tiny"``,
optionally ``?seed=N`` and ``&precision=exact`` (the fp32 parity mode).  Like the reference's builder (builder.py:33) the model name must contain
``llava`` or ``teochat``.  A directory path loads an HF-format checkpoint (merged, or a LoRA adapter over ``model_base``; a
separate image-tower checkpoint directory may be given through ``cache_dir``) via teochat_b200.checkpoint.
"""
from __future__ import annotations

import argparse
import json
import os
from pathlib import Path

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
    opts = dict(kv.partition("=")[::2] for kv in filter(None, query.split("&")))
    precision = opts.get("precision")                  # "exact": the fp32 parity mode (engine.TeoModel); default bf16
    if os.path.isdir(path):
        # HF-format directory: merged checkpoint, or LoRA adapter over `model_base` (builder.py:33-112)
        from .. import checkpoint as CK
        cfg = CK.read_config(path)
        sd = CK.load_state_dict(path, model_base if (model_base and os.path.isdir(str(model_base))) else None, tower_path=cache_dir)
        weights = TeoWeights.from_state_dict(sd, cfg, dev)
        del sd
        model = TeoModel(cfg, weights, dev, precision=precision)
        model.model.video_tower = None
        return CK.load_tokenizer(model_base if (model_base and os.path.isdir(str(model_base))) else path, cfg.llama.vocab_size), model, \
            TeoImageProcessor(cfg.vision.image_size)
    seed = int(opts.get("seed", 1234))
    cfg = TeoConfig.tiny() if model_name.endswith("-tiny") else TeoConfig.full()
    weights = TeoWeights.from_synthetic(cfg, seed, dev)
    model = TeoModel(cfg, weights, dev, precision=precision)
    model.model.video_tower = None                     # eval.py:31
    tokenizer = StubTokenizer(cfg.llama.vocab_size)
    processor = TeoImageProcessor(cfg.vision.image_size)   # processor['image'] (eval.py:33)
    return tokenizer, model, processor


def _load_split(dataset_name: str, data_cache_dir=None):
    """eval.py:152: ``load_dataset("jirvin16/TEOChatlas", split=f"eval_{hf_split}")`` — needs the `datasets` package and
    the hub (or a populated cache); raises with a clear message otherwise."""
    from .metrics import HF_SPLITS
    try:
        from datasets import load_dataset
    except ImportError as e:                                    # not in the offline image
        raise RuntimeError("the `datasets` package is not installed: pass `dataset=` (an iterable of TEOChatlas-style "
                           "examples) to eval()") from e
    return load_dataset("jirvin16/TEOChatlas", split=f"eval_{HF_SPLITS[dataset_name]}", cache_dir=data_cache_dir, trust_remote_code=True)


def eval(dataset_name, model_path, model_base, load_8bit=False, load_4bit=False, cache_dir=None, data_cache_dir=None,   # noqa: A001
         out_name=None, out_dir=None, prompt_strategy=None, chronological_prefix=True, conv_mode="v1", device="cuda",
         force_rerun=False, temperature=0.2, max_new_tokens=256, dataset=None, batch_size: int = 32):
    """Drop-in for ``videollava/eval/eval.py::eval`` (:37-171): same arguments, output-file naming, result caching
    (``force_rerun``), JSON layout and metric dispatch; returns the metrics dict it prints.

    Added: ``dataset`` (any iterable of TEOChatlas examples, so the driver runs without the hub) and ``batch_size``
    (examples per batched generate; the reference loops at bs=1).  Under torchrun every rank generates a contiguous
    shard of the examples with its own replica and the responses are gathered once at the end;
    rank 0 writes the file and scores it."""
    import torch.distributed as dist

    from .. import dist as TD
    from .inference import run_inference
    from .metrics import metrics_fn_for
    metrics_fn = metrics_fn_for(dataset_name)                   # ValueError for unknown datasets, before any work
    out_dir = Path("results") if out_dir is None else Path(out_dir)
    out_sub = out_dir / dataset_name
    if out_name is None:
        out_name = f"{get_model_name_from_path(str(model_path).partition('?')[0])}.json"
    if ".json" not in out_name:
        out_name = f"{out_name}.json"
    for arg, val in (("prompt_strategy", prompt_strategy), ("chronological_prefix", chronological_prefix)):
        if val is not None:
            out_name = out_name.replace(".json", f"_{arg}_{val}.json")
    out_path = out_sub / out_name
    rank, world, local = TD.init_from_env()
    if out_path.exists() and not force_rerun:
        print(f"Output file {out_path} already exists. Computing metrics without running inference.")
        with open(out_path) as f:
            outputs = json.load(f)
    else:
        if world > 1 and str(device).startswith("cuda"):
            device = f"cuda:{local}"
        tokenizer, model, processor = load_model(model_path, model_base, load_8bit=load_8bit, load_4bit=load_4bit,
                                                 cache_dir=cache_dir, device=device)
        examples = list(dataset if dataset is not None else _load_split(dataset_name, data_cache_dir))
        lo, hi = TD.shard_range(len(examples), rank, world)
        outputs = run_inference(examples[lo:hi], model, tokenizer, processor, prompt_strategy, chronological_prefix,
                                conv_mode, temperature, max_new_tokens, batch_size=batch_size)
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, outputs)              # strings/dicts: the one end-of-run exchange
            outputs = [o for part in parts for o in part]
        if rank == 0:
            out_sub.mkdir(parents=True, exist_ok=True)
            print(f"Saving outputs to {out_path}")
            with open(out_path, "w") as f:
                json.dump(outputs, f, indent=4)
    metrics = metrics_fn(outputs, dataset_name=dataset_name)
    if rank == 0:
        print(f"Metrics for dataset {dataset_name}:")
        for k, v in metrics.items():
            print(f"\t{k}: {v}")
    return metrics


def _str_or_none(v):
    return None if v == "" or v.lower() == "none" else v


def main(argv=None):
    """CLI of eval.py:180-199 (same flags)."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset_name", type=str, required=True)
    ap.add_argument("--model_path", type=str, required=True)
    ap.add_argument("--model_base", type=_str_or_none, default=None)
    ap.add_argument("--load_8bit", action="store_true")
    ap.add_argument("--load_4bit", action="store_true")
    ap.add_argument("--cache_dir", type=str, default=None)
    ap.add_argument("--data_cache_dir", type=str, default=None)
    ap.add_argument("--out_name", type=str, default=None)
    ap.add_argument("--out_dir", type=str, default=None)
    ap.add_argument("--prompt_strategy", type=str, default="interleave")
    ap.add_argument("--chronological_prefix", action="store_true")
    ap.add_argument("--device", type=str, default="cuda")
    ap.add_argument("--force_rerun", action="store_true")
    ap.add_argument("--temperature", type=float, default=0.2)
    ap.add_argument("--max_new_tokens", type=int, default=256)
    ap.add_argument("--batch_size", type=int, default=32)
    return eval(**vars(ap.parse_args(argv)))


if __name__ == "__main__":
    main()
