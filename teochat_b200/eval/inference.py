"""Drop-in for ``videollava/eval/inference.py``: same functions, same signatures, same string
results — ``run_inference_single`` (:23-77), ``replace_video_token`` (:11-20), ``extract_bboxes``
(:80-85), ``run_inference`` (:88-137) — plus ``run_inference_batch``, the batched/data-parallel
form the B200 build exists for (the reference loop is strictly bs=1).
"""
from __future__ import annotations

import re
from datetime import datetime
from typing import List, Optional, Sequence

import torch

from ..constants import DEFAULT_IMAGE_TOKEN, DEFAULT_VIDEO_TOKEN, IMAGE_TOKEN_INDEX
from ..conversation import SeparatorStyle, conv_templates
from ..mm_utils import KeywordsStoppingCriteria, tokenizer_image_token


def replace_video_token(prompt: str, image_paths, prompt_strategy):
    """inference.py:11-20"""
    if prompt_strategy is None:
        rep = DEFAULT_IMAGE_TOKEN * len(image_paths)
    elif prompt_strategy == "interleave":
        rep = "".join(f"Image {i + 1}: {DEFAULT_IMAGE_TOKEN}" for i in range(len(image_paths)))
    else:
        raise ValueError(f"Unknown prompt strategy: {prompt_strategy}")
    return prompt.replace(DEFAULT_VIDEO_TOKEN, rep)


def build_prompt(inp: str, image_paths, conv_mode="v1", timestamps=(), prompt_strategy="interleave",
                 chronological_prefix=True):
    """Prompt + frame order exactly as run_inference_single builds them (inference.py:37-55).
    Returns (prompt, ordered_image_paths, stop_str)."""
    conv = conv_templates[conv_mode].copy()
    conv.append_message(conv.roles[0], inp)
    conv.append_message(conv.roles[1], None)
    prompt = conv.get_prompt()
    if chronological_prefix:
        prompt = prompt.replace("times:", "times in chronological order:")
    image_paths = list(image_paths)
    if len(timestamps) > 0:
        order = sorted(zip(image_paths, timestamps), key=lambda t: datetime.strptime(t[1], "%Y-%m-%d"))
        image_paths = [p for p, _ in order]          # zip(*sorted(...)) in the reference
    prompt = replace_video_token(prompt, image_paths, prompt_strategy)
    stop_str = conv.sep if conv.sep_style != SeparatorStyle.TWO else conv.sep2
    return prompt, image_paths, stop_str


def run_inference_single(model, processor, tokenizer, inp, image_paths, conv_mode="v1", timestamps=[],
                         prompt_strategy="interleave", chronological_prefix=True, temperature=0.2,
                         max_new_tokens=256):
    """inference.py:23-77.  ``temperature > 0`` samples on the device (reference default 0.2);
    ``temperature <= 0`` selects greedy decoding (the parity path)."""
    prompt, image_paths, stop_str = build_prompt(inp, image_paths, conv_mode, timestamps, prompt_strategy,
                                                 chronological_prefix)
    if hasattr(processor, "preprocess_device") and getattr(model, "device", torch.device("cpu")).type == "cuda":
        tensors = list(processor.preprocess_device(list(image_paths), model.device))     # same chain, CUDA kernels
    else:                                                                                # a reference-style processor
        tensors = [processor.preprocess(i, return_tensors="pt")["pixel_values"][0] for i in image_paths]
        tensors = [t.to(model.device, dtype=torch.float32) for t in tensors]
    input_ids = tokenizer_image_token(prompt, tokenizer, IMAGE_TOKEN_INDEX, return_tensors="pt").unsqueeze(0).to(model.device)
    stopping_criteria = KeywordsStoppingCriteria([stop_str], tokenizer, input_ids)
    with torch.inference_mode():
        output_ids = model.generate(input_ids=input_ids, images=tensors, do_sample=temperature > 0,
                                    temperature=temperature, max_new_tokens=max_new_tokens, use_cache=True,
                                    stopping_criteria=[stopping_criteria])
    return tokenizer.decode(output_ids[0, input_ids.shape[1]:]).replace("</s>", "").strip()


def run_inference_batch(model, processor, tokenizer, inps: Sequence[str], image_paths_list: Sequence[Sequence],
                        conv_mode="v1", timestamps_list: Optional[Sequence[Sequence[str]]] = None,
                        prompt_strategy="interleave", chronological_prefix=True, temperature=0.0,
                        max_new_tokens=256, seed: int = 0) -> List[str]:
    """Batched greedy form of run_inference_single: one ViT pass over all frames, one ragged
    prefill, one graph-replayed decode loop.  Result i equals run_inference_single on example i."""
    ids_list, frames, stops = [], [], []
    on_device = hasattr(processor, "preprocess_device") and getattr(model, "device", torch.device("cpu")).type == "cuda"
    for n, (inp, paths) in enumerate(zip(inps, image_paths_list)):
        ts = timestamps_list[n] if timestamps_list is not None else ()
        prompt, paths, stop_str = build_prompt(inp, paths, conv_mode, ts, prompt_strategy, chronological_prefix)
        ids_list.append(tokenizer_image_token(prompt, tokenizer, IMAGE_TOKEN_INDEX))
        stops.append(stop_str)
        if on_device:        # raw uint8 over PCIe, ToTensor/Resize/CenterCrop/Normalize in CUDA (teo_resize_crop_normalize_u8)
            frames.append(processor.preprocess_device(list(paths), model.device))
        else:
            frames.append(torch.cat([processor.preprocess(p, return_tensors="pt")["pixel_values"] for p in paths]))
    outs = model.generate_batch(ids_list, pixel_values=frames, max_new_tokens=max_new_tokens,
                                temperature=temperature or 0.0, seed=seed)
    eos = model.cfg.llama.eos_token_id if hasattr(model, "cfg") else None
    for n, (ids, out) in enumerate(zip(ids_list, outs)):
        # templates whose stop string is not "</s>" (llava_llama_2: "<s>", plain: "\n"): the same KeywordsStoppingCriteria the
        # single-example path builds, evaluated on the host over the generated prefix
        prompt_ids = torch.tensor([ids])
        sc = KeywordsStoppingCriteria([stops[n]], tokenizer, prompt_ids)
        if eos is not None and sc.is_eos_only(eos):
            continue
        full = torch.cat([prompt_ids, torch.tensor([out], dtype=prompt_ids.dtype)], dim=1)
        for k in range(1, len(out) + 1):
            if sc(full[:, :len(ids) + k], None):
                outs[n] = out[:k]
                break
    return [tokenizer.decode(o).replace("</s>", "").strip() for o in outs]


def extract_bboxes(bbox_str):
    """inference.py:80-85"""
    pattern = re.compile(r"\[(\d+), (\d+), (\d+), (\d+)\]")
    return [list(map(int, m.groups())) for m in pattern.finditer(bbox_str)]


def run_inference(dataset, model, tokenizer, processor, prompt_strategy, chronological_prefix, conv_mode,
                  temperature, max_new_tokens, batch_size: int = 1):
    """inference.py:88-137 (same output dicts).  ``batch_size > 1`` routes greedy runs through
    run_inference_batch."""
    examples = list(dataset)
    responses: List[str] = []
    if batch_size > 1:
        for s in range(0, len(examples), batch_size):
            chunk = examples[s:s + batch_size]
            responses += run_inference_batch(model, processor, tokenizer,
                                             [e["conversations"][0]["value"] for e in chunk],
                                             [e["video"] for e in chunk], conv_mode=conv_mode,
                                             timestamps_list=[e["timestamp"] for e in chunk],
                                             prompt_strategy=prompt_strategy,
                                             chronological_prefix=chronological_prefix, temperature=temperature,
                                             max_new_tokens=max_new_tokens, seed=s)
    else:
        for e in examples:
            responses.append(run_inference_single(model, processor, tokenizer, e["conversations"][0]["value"], e["video"],
                                                  conv_mode=conv_mode, timestamps=e["timestamp"],
                                                  prompt_strategy=prompt_strategy,
                                                  chronological_prefix=chronological_prefix, temperature=temperature,
                                                  max_new_tokens=max_new_tokens))
    outputs = []
    for example, response in zip(examples, responses):
        output = {"response": response, "ground_truth": example["conversations"][1]["value"], "task": example["task"]}
        polygon = example.get("polygon", None)
        if polygon is not None:
            output["polygon"] = polygon
        input_bboxes = extract_bboxes(example["conversations"][0]["value"])
        output_bboxes = extract_bboxes(example["conversations"][1]["value"])
        if len(input_bboxes) > 0:
            output["input_bboxes"] = input_bboxes
        if len(output_bboxes) > 0:
            output["output_bboxes"] = output_bboxes
        outputs.append(output)
    return outputs
