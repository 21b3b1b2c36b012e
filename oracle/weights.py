"""ORACLE — test infrastructure only.  Synthetic ("random-init") checkpoint in HF state-dict
names (SURVEY.md §5 checkpoint row): CLIP-L tower, mlp2x_gelu projector, LLaMA-2-7B.

Std per tensor follows the reference's initialisers (modeling_image.py:179-230 for CLIP;
nn.Linear default for the projector, multimodal_projector/builder.py:41-48; HF Llama
N(0, 0.02)).  Deviation, stated: biases and norm gains — which those initialisers set to 0 / 1 —
are drawn N(0,0.02) / N(1,0.02) so that parity tests exercise every parameter.
Values come from oracle/hashinit (counter-based, bit-reproducible on any machine).
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, Tuple

import torch

from . import hashinit

VIT = "model.image_tower.image_tower."


def tensor_specs(cfg) -> Iterator[Tuple[str, tuple, float, float]]:
    """Yield (name, shape, std, mean) for every parameter of the hot path."""
    v, l = cfg.vision, cfg.llama
    d, L = v.hidden_size, v.num_hidden_layers
    f = v.initializer_factor
    yield VIT + "embeddings.class_embedding", (d,), d ** -0.5 * f, 0.0
    yield VIT + "embeddings.patch_embedding.weight", (d, v.num_channels, v.patch_size, v.patch_size), v.initializer_range * f, 0.0
    yield VIT + "embeddings.position_embedding.weight", (v.num_positions, d), v.initializer_range * f, 0.0
    yield VIT + "pre_layrnorm.weight", (d,), 0.02, 1.0
    yield VIT + "pre_layrnorm.bias", (d,), 0.02, 0.0
    in_std = d ** -0.5 * (2 * L) ** -0.5 * f
    out_std = d ** -0.5 * f
    fc_std = (2 * d) ** -0.5 * f
    for i in range(L):
        p = f"{VIT}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj"):
            yield p + f"self_attn.{n}.weight", (d, d), in_std, 0.0
            yield p + f"self_attn.{n}.bias", (d,), 0.02, 0.0
        yield p + "self_attn.out_proj.weight", (d, d), out_std, 0.0
        yield p + "self_attn.out_proj.bias", (d,), 0.02, 0.0
        yield p + "layer_norm1.weight", (d,), 0.02, 1.0
        yield p + "layer_norm1.bias", (d,), 0.02, 0.0
        yield p + "mlp.fc1.weight", (v.intermediate_size, d), fc_std, 0.0
        yield p + "mlp.fc1.bias", (v.intermediate_size,), 0.02, 0.0
        yield p + "mlp.fc2.weight", (d, v.intermediate_size), in_std, 0.0
        yield p + "mlp.fc2.bias", (d,), 0.02, 0.0
        yield p + "layer_norm2.weight", (d,), 0.02, 1.0
        yield p + "layer_norm2.bias", (d,), 0.02, 0.0
    h = l.hidden_size
    yield "model.mm_projector.0.weight", (h, d), (3.0 * d) ** -0.5, 0.0
    yield "model.mm_projector.0.bias", (h,), 0.02, 0.0
    yield "model.mm_projector.2.weight", (h, h), (3.0 * h) ** -0.5, 0.0
    yield "model.mm_projector.2.bias", (h,), 0.02, 0.0
    s = l.initializer_range
    yield "model.embed_tokens.weight", (l.vocab_size, h), s, 0.0
    for i in range(l.num_hidden_layers):
        p = f"model.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            yield p + f"self_attn.{n}.weight", (h, h), s, 0.0
        yield p + "mlp.gate_proj.weight", (l.intermediate_size, h), s, 0.0
        yield p + "mlp.up_proj.weight", (l.intermediate_size, h), s, 0.0
        yield p + "mlp.down_proj.weight", (h, l.intermediate_size), s, 0.0
        yield p + "input_layernorm.weight", (h,), 0.02, 1.0
        yield p + "post_attention_layernorm.weight", (h,), 0.02, 1.0
    yield "model.norm.weight", (h,), 0.02, 1.0
    yield "lm_head.weight", (l.vocab_size, h), s, 0.0


def make_state_dict(cfg, seed: int, dtype=torch.float32, bf16_values: bool = True,
                    prefixes=None) -> Dict[str, torch.Tensor]:
    """Generate the synthetic checkpoint.  ``bf16_values`` rounds every value to bf16 (the
    storage type of the B200 build) before returning it in ``dtype``; ``prefixes`` restricts the
    generated tensors (e.g. only the tower)."""
    sd = {}
    for name, shape, std, mean in tensor_specs(cfg):
        if prefixes is not None and not any(name.startswith(p) for p in prefixes):
            continue
        t = hashinit.hash_normal(shape, hashinit.tensor_seed(seed, name), std, mean)
        if bf16_values:
            t = t.to(torch.bfloat16)
        sd[name] = t.to(dtype)
    return sd


def synthetic_frames_u8(n_frames: int, size: int, seed: int) -> torch.Tensor:
    """u8 [n,H,W,3] NHWC uniform{0..255} frames (SURVEY.md §8d synthetic inputs)."""
    return hashinit.hash_u8((n_frames, size, size, 3), hashinit.tensor_seed(seed, "frames"))
