"""HF-format checkpoint loading for the hot path: merged checkpoints, LoRA
checkpoints (adapter + ``non_lora_trainables.bin``) over a base model, and a separate image-tower
checkpoint — the branches of ``videollava/model/builder.py:33-155`` that matter for TEOChat
(``scripts/merge_lora_weights.py:10-31`` for the merge semantics).  The image tower's encoder is PEFT-wrapped by the
reference itself (``modeling_image.py:773-792``, default ``lora_r=2, lora_alpha=16``): its keys look like
``…encoder.base_model.model.layers.N.self_attn.q_proj.base_layer.weight`` + ``…lora_A/B.default.weight`` and the adapter
runs UNMERGED at scaling 8 — here it is merged into the tower weights at load (``tower_lora_scaling``).  Everything here is host-side tensor plumbing; the result is an HF-named state dict that
``TeoWeights.from_state_dict`` lays out for the kernels.

No checkpoint, tokenizer or network exists in the build container, so this module is exercised on synthetic
directories written in the same formats (tests/test_checkpoint_cpu.py); it has not met the released weights.
"""
from __future__ import annotations

import glob
import json
import os
import re
from typing import Dict, Iterator, Optional, Tuple

import torch

from .config import LlamaConfig, TeoConfig, VisionConfig

VIT = "model.image_tower.image_tower."


def _get(d: dict, key: str, default):
    v = d.get(key, default)
    return default if v is None else v


def read_config(model_dir: str, tower_dir: Optional[str] = None) -> TeoConfig:
    """config.json → TeoConfig.  The attributes are the ones the reference reads with getattr
    (llava_arch.py:32-37,296,312; builder.py:139-149)."""
    with open(os.path.join(model_dir, "config.json")) as f:
        c = json.load(f)
    l = LlamaConfig(hidden_size=c["hidden_size"], intermediate_size=c["intermediate_size"],
                    num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
                    vocab_size=c["vocab_size"], rms_norm_eps=_get(c, "rms_norm_eps", 1e-5),
                    rope_theta=_get(c, "rope_theta", 10000.0), max_position_embeddings=_get(c, "max_position_embeddings", 4096),
                    bos_token_id=_get(c, "bos_token_id", 1), eos_token_id=_get(c, "eos_token_id", 2))
    if _get(c, "num_key_value_heads", l.num_attention_heads) != l.num_attention_heads:
        raise NotImplementedError("grouped-query attention is not part of the LLaMA-2-7B path")
    vc = {}
    tdir = tower_dir or _get(c, "mm_image_tower", None)
    if tdir and os.path.isfile(os.path.join(tdir, "config.json")):
        with open(os.path.join(tdir, "config.json")) as f:
            tc = json.load(f)
        vc = tc.get("vision_config", tc)
    elif "vision_config" in c:
        vc = c["vision_config"]
    d = VisionConfig()
    v = VisionConfig(hidden_size=_get(vc, "hidden_size", d.hidden_size), intermediate_size=_get(vc, "intermediate_size", d.intermediate_size),
                     num_hidden_layers=_get(vc, "num_hidden_layers", d.num_hidden_layers),
                     num_attention_heads=_get(vc, "num_attention_heads", d.num_attention_heads),
                     image_size=_get(vc, "image_size", d.image_size), patch_size=_get(vc, "patch_size", d.patch_size),
                     layer_norm_eps=_get(vc, "layer_norm_eps", d.layer_norm_eps), hidden_act=_get(vc, "hidden_act", d.hidden_act))
    if _get(vc, "add_time_attn", False):
        raise NotImplementedError("temporal attention in the image tower (add_time_attn) is not used by TEOChat")
    cfg = TeoConfig(vision=v, llama=l, mm_projector_type=_get(c, "mm_projector_type", "mlp2x_gelu"),
                    mm_vision_select_layer=_get(c, "mm_vision_select_layer", -2),
                    mm_vision_select_feature=_get(c, "mm_vision_select_feature", "patch"),
                    mm_use_im_start_end=_get(c, "mm_use_im_start_end", False), mm_use_im_patch_token=_get(c, "mm_use_im_patch_token", False),
                    tokenizer_model_max_length=c.get("tokenizer_model_max_length"), tokenizer_padding_side=_get(c, "tokenizer_padding_side", "right"))
    if _get(c, "mm_hidden_size", v.hidden_size) != v.hidden_size:
        raise ValueError(f"mm_hidden_size {c['mm_hidden_size']} != tower hidden size {v.hidden_size}")
    return cfg


def iter_shards(model_dir: str) -> Iterator[Dict[str, torch.Tensor]]:
    """Every weight shard of a directory: *.safetensors first, else pytorch_model*.bin."""
    st = sorted(p for p in glob.glob(os.path.join(model_dir, "*.safetensors")) if not os.path.basename(p).startswith("adapter_"))
    if st:
        from safetensors.torch import load_file
        for p in st:
            yield load_file(p)
        return
    bins = sorted(glob.glob(os.path.join(model_dir, "pytorch_model*.bin")))
    if not bins:
        raise FileNotFoundError(f"no *.safetensors or pytorch_model*.bin under {model_dir}")
    for p in bins:
        yield torch.load(p, map_location="cpu", weights_only=True)


def _canon(name: str) -> Optional[str]:
    """Map the spellings found in LLaVA / LanguageBind / PEFT checkpoints onto the HF names of param_specs."""
    for pre in ("base_model.model.", "base_model."):
        if name.startswith(pre):
            name = name[len(pre):]
    if name.startswith("model.model."):
        name = name[len("model."):]
    name = name.replace(".base_layer.", ".")                       # PEFT-wrapped Linear
    # the image tower's ENCODER is itself a PeftModel (LanguageBindImage.convert_to_lora, modeling_image.py:773-792:
    # vision_model.encoder = get_peft_model(encoder, …)), so its keys carry an inner wrapper prefix
    name = name.replace(".encoder.base_model.model.", ".encoder.")
    if name.startswith("vision_model."):                          # stand-alone LanguageBind_Image / CLIPVisionModel checkpoint
        name = VIT + name[len("vision_model."):]
    if name.startswith("model.image_tower.image_tower.vision_model."):
        name = VIT + name[len("model.image_tower.image_tower.vision_model."):]
    return name


def _lora_key(name: str) -> Optional[Tuple[str, str]]:
    m = re.match(r"(.*)\.lora_([AB])(?:\.default)?\.weight$", name)
    return (m.group(1) + ".weight", m.group(2)) if m else None


def merge_lora_(sd: Dict[str, torch.Tensor], adapter: Dict[str, torch.Tensor], scaling: float) -> int:
    """W += scaling · B @ A for every (lora_A, lora_B) pair, in fp32 (PEFT merge_and_unload semantics). Returns #merged."""
    pairs: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, t in adapter.items():
        lk = _lora_key(_canon(k))
        if lk:
            pairs.setdefault(lk[0], {})[lk[1]] = t
    n = 0
    for wname, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise ValueError(f"incomplete LoRA pair for {wname}")
        if wname not in sd:
            raise KeyError(f"LoRA target {wname} is not in the base checkpoint")
        w = sd[wname].to(torch.float32)
        w += scaling * (ab["B"].to(torch.float32) @ ab["A"].to(torch.float32))
        sd[wname] = w
        n += 1
    return n


def tower_lora_scaling(*dirs: Optional[str]) -> float:
    """lora_alpha / lora_r of the image tower's built-in adapter: the first config.json among `dirs` that names them
    (top level or under vision_config), else the class defaults of configuration_image.py:200-201 (r=2, alpha=16)."""
    for d in dirs:
        if not d or not os.path.isfile(os.path.join(d, "config.json")):
            continue
        with open(os.path.join(d, "config.json")) as f:
            c = json.load(f)
        for scope in (c.get("vision_config") or {}, c):
            if "lora_r" in scope:
                r = float(scope["lora_r"])
                if r == 0:
                    return 0.0
                return float(scope.get("lora_alpha", 16)) / r
    return 16.0 / 2.0


def load_state_dict(model_path: str, model_base: Optional[str] = None, tower_path: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """HF-named state dict of the hot path from
       - a merged checkpoint directory (model_base None), or
       - a LoRA directory (adapter_model.* + adapter_config.json [+ non_lora_trainables.bin]) over `model_base`,
       plus, if the tower weights are not inside, a separate tower checkpoint directory."""
    base_dir = model_base or model_path
    sd: Dict[str, torch.Tensor] = {}
    tower_lora: Dict[str, torch.Tensor] = {}     # the tower's own (unmerged) adapter: it runs at lora_alpha/lora_r in the reference
    for shard in iter_shards(base_dir):
        for k, t in shard.items():
            ck = _canon(k)
            if _lora_key(ck) is None:
                sd[ck] = t
            elif ck.startswith(VIT):
                tower_lora[ck] = t
    tower_src = base_dir
    if tower_path and not any(k.startswith(VIT) for k in sd):
        tower_src = tower_path
        for shard in iter_shards(tower_path):
            for k, t in shard.items():
                ck = _canon(k)
                if not ck.startswith(VIT):
                    continue
                if _lora_key(ck) is None:
                    sd[ck] = t
                else:
                    tower_lora[ck] = t
    if model_base is not None:
        nl = os.path.join(model_path, "non_lora_trainables.bin")          # projector etc. trained without LoRA (builder.py:52-66)
        if os.path.exists(nl):
            for k, t in torch.load(nl, map_location="cpu", weights_only=True).items():
                sd[_canon(k)] = t
        with open(os.path.join(model_path, "adapter_config.json")) as f:
            ac = json.load(f)
        scaling = float(ac["lora_alpha"]) / float(ac["r"])
        adapters = glob.glob(os.path.join(model_path, "adapter_model.safetensors")) or glob.glob(os.path.join(model_path, "adapter_model.bin"))
        if not adapters:
            raise FileNotFoundError(f"no adapter_model.* under {model_path}")
        if adapters[0].endswith(".safetensors"):
            from safetensors.torch import load_file
            ad = load_file(adapters[0])
        else:
            ad = torch.load(adapters[0], map_location="cpu", weights_only=True)
        merge_lora_(sd, ad, scaling)
        # a LoRA checkpoint may carry re-trained tower adapters among its non-LoRA trainables: they win over the tower's own
        for k in [k for k in sd if _lora_key(k) is not None]:
            t = sd.pop(k)
            if k.startswith(VIT):
                tower_lora[k] = t
    if tower_lora:
        s = tower_lora_scaling(tower_src, model_path, model_base)
        if s != 0.0:
            merge_lora_(sd, tower_lora, s)
    return sd


def load_tokenizer(model_path: str, vocab_size: int):
    """The reference uses the slow sentencepiece LLaMA tokenizer (builder.py:111); fall back to the offline stub."""
    if os.path.exists(os.path.join(model_path, "tokenizer.model")):
        from transformers import AutoTokenizer
        return AutoTokenizer.from_pretrained(model_path, use_fast=False)
    from .tokenizer import StubTokenizer
    return StubTokenizer(vocab_size)
