"""ORACLE — test infrastructure only.  CPU restatement of the hot path's arithmetic.

Follows the reference call order and the HF-4.31 op order it executes
(transformers==4.31.0 is pinned at pyproject.toml:16 and is not
vendored, so its published algorithm is restated here):

  * CLIP vision tower  — modeling_image.py:610-672 (CLIPVisionTransformer.forward),
    :136-151 (CLIPEncoderLayer spatial path), HF CLIPVisionEmbeddings/CLIPAttention/CLIPMLP;
    feature select languagebind/__init__.py:121-129 (hidden_states[-2], drop CLS).
  * projector          — multimodal_projector/builder.py:41-48 (Linear, GELU(erf), Linear).
  * splice             — llava_arch.py:251-331.
  * LLaMA              — HF-4.31 LlamaModel (pre-RMSNorm, rotate-half RoPE, causal MHA with
    fp32 softmax, SwiGLU), lm_head, logits.float() (llava_llama.py:56-99).
  * greedy decode      — argmax(logits[:, -1]).

Arithmetic is fp32.  ``policy="bf16"`` additionally rounds activations to bf16 at the points
where the B200 build stores them to HBM (a subset of the points where an HF bf16 model
rounds), so token-id parity is not dominated by storage rounding; ``policy="fp32"`` is the
unrounded mathematical reference the 1e-2 logit tolerance is quoted against.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200            # constants.py:9
VIT = "model.image_tower.image_tower."
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # processing_image.py:7
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)   # processing_image.py:8


def _rounder(policy: str):
    if policy == "fp32":
        return lambda x: x
    if policy == "bf16":
        return lambda x: x.to(torch.bfloat16).to(torch.float32)
    raise ValueError(policy)


def normalize_u8_nhwc(frames_u8: torch.Tensor) -> torch.Tensor:
    """u8 [N,H,W,3] → f32 [N,3,H,W]: ToTensor + Normalize (processing_image.py:18,22).
    Written as (x/255 - mean)/std exactly like torchvision does it."""
    x = frames_u8.permute(0, 3, 1, 2).to(torch.float32) / 255.0
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32)[None, :, None, None]
    std = torch.tensor(CLIP_STD, dtype=torch.float32)[None, :, None, None]
    return (x - mean) / std


def _act(name: str, x: torch.Tensor) -> torch.Tensor:
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    if name == "gelu":
        return F.gelu(x)
    raise ValueError(f"unknown activation {name}")


def vit_hidden_states(sd: Dict[str, torch.Tensor], cfg, pixel_values: torch.Tensor,
                      policy: str = "fp32", n_layers: Optional[int] = None) -> List[torch.Tensor]:
    """All hidden states [embeddings(after pre_layrnorm), layer1, ...] as HF returns them
    (modeling_image.py:399-431: hidden_states[0] is the *post*-pre_layrnorm input)."""
    r = _rounder(policy)
    v = cfg.vision
    d, H, hd = v.hidden_size, v.num_attention_heads, v.head_dim
    N = pixel_values.shape[0]
    x = r(pixel_values.to(torch.float32))
    w = sd[VIT + "embeddings.patch_embedding.weight"].float()
    patches = F.conv2d(x, w, stride=v.patch_size)                      # [N,d,g,g], no bias
    patches = r(patches.flatten(2).transpose(1, 2))                     # [N,g*g,d]
    cls = sd[VIT + "embeddings.class_embedding"].float().expand(N, 1, d)
    h = torch.cat([cls, patches], dim=1) + sd[VIT + "embeddings.position_embedding.weight"].float()[None]
    h = r(F.layer_norm(h, (d,), sd[VIT + "pre_layrnorm.weight"].float(),
                       sd[VIT + "pre_layrnorm.bias"].float(), v.layer_norm_eps))
    states = [h]
    L = v.num_hidden_layers if n_layers is None else n_layers
    scale = hd ** -0.5
    for i in range(L):
        p = f"{VIT}encoder.layers.{i}."
        y = r(F.layer_norm(h, (d,), sd[p + "layer_norm1.weight"].float(), sd[p + "layer_norm1.bias"].float(),
                           v.layer_norm_eps))
        q = r(F.linear(y, sd[p + "self_attn.q_proj.weight"].float(), sd[p + "self_attn.q_proj.bias"].float()))
        k = r(F.linear(y, sd[p + "self_attn.k_proj.weight"].float(), sd[p + "self_attn.k_proj.bias"].float()))
        vv = r(F.linear(y, sd[p + "self_attn.v_proj.weight"].float(), sd[p + "self_attn.v_proj.bias"].float()))
        q = q.view(N, -1, H, hd).transpose(1, 2) * scale               # HF-4.31 CLIPAttention scales q
        k = k.view(N, -1, H, hd).transpose(1, 2)
        vv = vv.view(N, -1, H, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        m = s.amax(dim=-1, keepdim=True)
        pexp = torch.exp(s - m)
        o = (r(pexp) @ vv) / pexp.sum(dim=-1, keepdim=True)            # flash-style normalisation
        o = r(o.transpose(1, 2).reshape(N, -1, d))
        h = r(h + F.linear(o, sd[p + "self_attn.out_proj.weight"].float(), sd[p + "self_attn.out_proj.bias"].float()))
        y = r(F.layer_norm(h, (d,), sd[p + "layer_norm2.weight"].float(), sd[p + "layer_norm2.bias"].float(),
                           v.layer_norm_eps))
        y = r(_act(v.hidden_act, F.linear(y, sd[p + "mlp.fc1.weight"].float(), sd[p + "mlp.fc1.bias"].float())))
        h = r(h + F.linear(y, sd[p + "mlp.fc2.weight"].float(), sd[p + "mlp.fc2.bias"].float()))
        states.append(h)
    return states


def vit_features(sd, cfg, pixel_values, policy="fp32") -> torch.Tensor:
    """LanguageBindImageTower.forward + feature_select (languagebind/__init__.py:121-146)."""
    n_run = cfg.vit_layers_run
    hs = vit_hidden_states(sd, cfg, pixel_values, policy, n_layers=n_run)[n_run]
    if cfg.mm_vision_select_feature == "patch":
        return hs[:, 1:]
    if cfg.mm_vision_select_feature == "cls_patch":
        return hs
    raise ValueError(f"Unexpected select feature: {cfg.mm_vision_select_feature}")


def projector(sd, cfg, feats: torch.Tensor, policy="fp32") -> torch.Tensor:
    """mlp2x_gelu (multimodal_projector/builder.py:41-48)."""
    if cfg.mm_projector_type != "mlp2x_gelu":
        raise ValueError(f"Unknown projector type: {cfg.mm_projector_type}")
    r = _rounder(policy)
    y = r(F.gelu(F.linear(feats, sd["model.mm_projector.0.weight"].float(), sd["model.mm_projector.0.bias"].float())))
    return r(F.linear(y, sd["model.mm_projector.2.weight"].float(), sd["model.mm_projector.2.bias"].float()))


def encode_images(sd, cfg, pixel_values, policy="fp32") -> torch.Tensor:
    """llava_arch.py:137-140."""
    return projector(sd, cfg, vit_features(sd, cfg, pixel_values, policy), policy)


def splice(sd, cfg, input_ids: Sequence[int], image_features: torch.Tensor) -> torch.Tensor:
    """One sample of prepare_inputs_labels_for_multimodal (llava_arch.py:251-299): text chunks
    between IMAGE_TOKEN_INDEX positions are embedded and interleaved with per-image feature
    blocks; the result is truncated to tokenizer_model_max_length if the config sets it."""
    ids = torch.as_tensor(list(input_ids), dtype=torch.long)
    E = sd["model.embed_tokens.weight"].float()
    pos = [-1] + torch.where(ids == IMAGE_TOKEN_INDEX)[0].tolist() + [ids.shape[0]]
    n_img = len(pos) - 2
    if n_img > image_features.shape[0]:
        raise IndexError("more <image> tokens than image features")
    parts = []
    for i in range(len(pos) - 1):
        parts.append(E[ids[pos[i] + 1:pos[i + 1]]])
        if i < n_img:
            parts.append(image_features[i].float())
    out = torch.cat(parts, dim=0)
    maxlen = getattr(cfg, "tokenizer_model_max_length", None)
    if maxlen is not None:
        out = out[:maxlen]
    return out


def _rope_tables(positions: torch.Tensor, hd: int, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = positions.to(torch.float32)[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


class LlamaOracle:
    """Single-sequence LLaMA with a growing KV cache (HF-4.31 LlamaModel op order)."""

    def __init__(self, sd, cfg, policy="fp32"):
        self.sd, self.cfg, self.l = sd, cfg, cfg.llama
        self.policy = policy
        self.r = _rounder(policy)
        self.k: List[Optional[torch.Tensor]] = [None] * self.l.num_hidden_layers
        self.v: List[Optional[torch.Tensor]] = [None] * self.l.num_hidden_layers
        self.len = 0

    def _rms(self, x, w):
        var = x.pow(2).mean(-1, keepdim=True)
        return self.r(x * torch.rsqrt(var + self.l.rms_norm_eps) * w.float())

    def forward(self, x: torch.Tensor, last_only: bool = True) -> torch.Tensor:
        """x: [S,h] new input embeddings at positions [len, len+S) → logits f32 [S or 1, vocab]."""
        l, r, sd = self.l, self.r, self.sd
        S, H, hd = x.shape[0], l.num_attention_heads, l.head_dim
        pos = torch.arange(self.len, self.len + S)
        cos, sin = _rope_tables(pos, hd, l.rope_theta)
        x = r(x.float())
        for i in range(l.num_hidden_layers):
            p = f"model.layers.{i}."
            y = self._rms(x, sd[p + "input_layernorm.weight"])
            q = r(F.linear(y, sd[p + "self_attn.q_proj.weight"].float())).view(S, H, hd).transpose(0, 1)
            k = r(F.linear(y, sd[p + "self_attn.k_proj.weight"].float())).view(S, H, hd).transpose(0, 1)
            v = r(F.linear(y, sd[p + "self_attn.v_proj.weight"].float())).view(S, H, hd).transpose(0, 1)
            q = r(q * cos[None] + _rotate_half(q) * sin[None])
            k = r(k * cos[None] + _rotate_half(k) * sin[None])
            self.k[i] = k if self.k[i] is None else torch.cat([self.k[i], k], dim=1)
            self.v[i] = v if self.v[i] is None else torch.cat([self.v[i], v], dim=1)
            K, V = self.k[i], self.v[i]
            s = (q @ K.transpose(-1, -2)) / math.sqrt(hd)              # HF-4.31: divide after QK^T
            T = K.shape[1]
            causal = torch.arange(T)[None, :] <= pos[:, None]
            s = s.masked_fill(~causal[None], float("-inf"))
            m = s.amax(dim=-1, keepdim=True)
            pexp = torch.exp(s - m)
            o = (r(pexp) @ V) / pexp.sum(dim=-1, keepdim=True)
            o = r(o.transpose(0, 1).reshape(S, H * hd))
            x = r(x + F.linear(o, sd[p + "self_attn.o_proj.weight"].float()))
            y = self._rms(x, sd[p + "post_attention_layernorm.weight"])
            g = r(F.linear(y, sd[p + "mlp.gate_proj.weight"].float()))
            u = r(F.linear(y, sd[p + "mlp.up_proj.weight"].float()))
            a = r(F.silu(g) * u)
            x = r(x + F.linear(a, sd[p + "mlp.down_proj.weight"].float()))
        self.len += S
        if last_only:
            x = x[-1:]
        y = self._rms(x, sd["model.norm.weight"])
        return F.linear(y, sd["lm_head.weight"].float())


def generate_greedy(sd, cfg, input_ids: Sequence[int], pixel_values: torch.Tensor, max_new_tokens: int,
                    policy: str = "fp32", eos_token_id: Optional[int] = 2, return_logits: bool = False):
    """Reference run_inference_single's model.generate with greedy decoding
    (inference.py:63-72 with do_sample=False).  Returns the new token ids (and per-step
    last-position logits).  Stops after emitting eos like KeywordsStoppingCriteria(["</s>"])."""
    feats = encode_images(sd, cfg, pixel_values, policy)
    emb = splice(sd, cfg, input_ids, feats)
    lm = LlamaOracle(sd, cfg, policy)
    logits = lm.forward(emb)
    out, all_logits = [], []
    E = sd["model.embed_tokens.weight"].float()
    for _ in range(max_new_tokens):
        all_logits.append(logits[0].clone())
        tok = int(torch.argmax(logits[0]))
        out.append(tok)
        if eos_token_id is not None and tok == eos_token_id:
            break
        if len(out) == max_new_tokens:
            break
        logits = lm.forward(E[tok][None])
    if return_logits:
        return out, torch.stack(all_logits)
    return out
