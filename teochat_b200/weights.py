"""Device-resident parameters of the hot path in the fused layouts the kernels read.

HF state-dict names: ``model.image_tower.image_tower.…`` (CLIP-L tower),
``model.mm_projector.{0,2}.{weight,bias}``, ``model.embed_tokens.weight``,
``model.layers.N.…``, ``model.norm.weight``, ``lm_head.weight``.

Fused layouts (all bf16, row-major, K contiguous — the K-major operands of the tcgen05 GEMM):
  ViT layer   qkv_w [3d,d] = q;k;v rows stacked, qkv_b [3d];   patch_w [d, kpad] (im2col order,
              zero padded from 3·14·14 = 588 to kpad = 640 so K is a whole number of 64-wide tiles)
  LLaMA layer qkv_w [3h,h] = q;k;v;   gate_up_w [2I,h] = gate;up while loading, then (I % 32 == 0) interleaved in
  blocks of 32 rows | gate 32 | up 32 | so the GEMM epilogue can apply SwiGLU (TEO_ACT_SWIGLU_PAIRS)

Two sources: ``from_synthetic`` (counter-based pseudo-normal init generated ON THE GPU by
teo_init_normal_hash_bf16; std per tensor follows the reference's initialisers,
modeling_image.py:179-230 / nn.Linear default / HF Llama N(0,0.02)) and ``from_state_dict``
(any HF-named dict of tensors, e.g. a real checkpoint shard or the test oracle's dict).
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, Iterator, Optional, Tuple

import torch

from . import lib as L
from .config import TeoConfig

VIT = "model.image_tower.image_tower."
_GOLDEN = 0x9E3779B97F4A7C15
_MASK = 0xFFFFFFFFFFFFFFFF
_IH4_STD = math.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0)


def _fnv1a64(name: str) -> int:
    h = 0xCBF29CE484222325
    for b in name.encode("utf-8"):
        h = ((h ^ b) * 0x100000001B3) & _MASK
    return h


def tensor_seed(global_seed: int, name: str) -> int:
    return (_fnv1a64(name) ^ ((global_seed * _GOLDEN) & _MASK)) & _MASK


def hash_scale(std: float) -> float:
    """fp32 multiplier that maps the centred Irwin-Hall(4) integer to the requested std."""
    return float(torch.tensor(std / _IH4_STD, dtype=torch.float64).to(torch.float32))


def kpad_for(cfg: TeoConfig) -> int:
    return (cfg.vision.patch_dim + 63) // 64 * 64


def param_specs(cfg: TeoConfig) -> Iterator[Tuple[str, tuple, float, float]]:
    """(HF name, shape, std, mean) of every parameter on the hot path."""
    v, l = cfg.vision, cfg.llama
    d, nl, fac = v.hidden_size, v.num_hidden_layers, v.initializer_factor
    yield VIT + "embeddings.class_embedding", (d,), d ** -0.5 * fac, 0.0
    yield VIT + "embeddings.patch_embedding.weight", (d, v.num_channels, v.patch_size, v.patch_size), v.initializer_range * fac, 0.0
    yield VIT + "embeddings.position_embedding.weight", (v.num_positions, d), v.initializer_range * fac, 0.0
    yield VIT + "pre_layrnorm.weight", (d,), 0.02, 1.0
    yield VIT + "pre_layrnorm.bias", (d,), 0.02, 0.0
    attn_in = d ** -0.5 * (2 * nl) ** -0.5 * fac
    for i in range(nl):
        p = f"{VIT}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj"):
            yield p + f"self_attn.{n}.weight", (d, d), attn_in, 0.0
            yield p + f"self_attn.{n}.bias", (d,), 0.02, 0.0
        yield p + "self_attn.out_proj.weight", (d, d), d ** -0.5 * fac, 0.0
        yield p + "self_attn.out_proj.bias", (d,), 0.02, 0.0
        yield p + "layer_norm1.weight", (d,), 0.02, 1.0
        yield p + "layer_norm1.bias", (d,), 0.02, 0.0
        yield p + "mlp.fc1.weight", (v.intermediate_size, d), (2 * d) ** -0.5 * fac, 0.0
        yield p + "mlp.fc1.bias", (v.intermediate_size,), 0.02, 0.0
        yield p + "mlp.fc2.weight", (d, v.intermediate_size), attn_in, 0.0
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


class TeoWeights:
    """Fused device buffers + a map HF name → (buffer key, row slice) used by both loaders."""

    def __init__(self, cfg: TeoConfig, device: torch.device):
        self.cfg, self.device = cfg, device
        v, l = cfg.vision, cfg.llama
        d, h, I = v.hidden_size, l.hidden_size, l.intermediate_size
        self.kpad = kpad_for(cfg)
        bf = dict(dtype=torch.bfloat16, device=device)
        t: Dict[str, torch.Tensor] = {}
        t["vit.patch_w"] = torch.zeros(d, self.kpad, **bf)
        t["vit.cls"] = torch.empty(d, **bf)
        t["vit.pos"] = torch.empty(v.num_positions, d, **bf)
        t["vit.pre_ln_w"] = torch.empty(d, **bf)
        t["vit.pre_ln_b"] = torch.empty(d, **bf)
        for i in range(v.num_hidden_layers):
            p = f"vit.{i}."
            t[p + "ln1_w"], t[p + "ln1_b"] = torch.empty(d, **bf), torch.empty(d, **bf)
            t[p + "qkv_w"], t[p + "qkv_b"] = torch.empty(3 * d, d, **bf), torch.empty(3 * d, **bf)
            t[p + "out_w"], t[p + "out_b"] = torch.empty(d, d, **bf), torch.empty(d, **bf)
            t[p + "ln2_w"], t[p + "ln2_b"] = torch.empty(d, **bf), torch.empty(d, **bf)
            t[p + "fc1_w"], t[p + "fc1_b"] = torch.empty(v.intermediate_size, d, **bf), torch.empty(v.intermediate_size, **bf)
            t[p + "fc2_w"], t[p + "fc2_b"] = torch.empty(d, v.intermediate_size, **bf), torch.empty(d, **bf)
        t["proj.w0"], t["proj.b0"] = torch.empty(h, d, **bf), torch.empty(h, **bf)
        t["proj.w2"], t["proj.b2"] = torch.empty(h, h, **bf), torch.empty(h, **bf)
        t["llama.embed"] = torch.empty(l.vocab_size, h, **bf)
        for i in range(l.num_hidden_layers):
            p = f"llama.{i}."
            t[p + "in_norm"], t[p + "post_norm"] = torch.empty(h, **bf), torch.empty(h, **bf)
            t[p + "qkv_w"] = torch.empty(3 * h, h, **bf)
            t[p + "o_w"] = torch.empty(h, h, **bf)
            t[p + "gate_up_w"] = torch.empty(2 * I, h, **bf)
            t[p + "down_w"] = torch.empty(h, I, **bf)
        t["llama.final_norm"] = torch.empty(h, **bf)
        t["llama.lm_head"] = torch.empty(l.vocab_size, h, **bf)
        self.t = t
        self.blocked: Dict[str, bool] = {}
        self.gate_up_interleaved = False
        self.ln_folded = False

    # HF name → destination view (contiguous row slice of a fused buffer), or None for patch_w
    def _dest(self, name: str) -> Optional[torch.Tensor]:
        t, cfg = self.t, self.cfg
        d, h, I = cfg.vision.hidden_size, cfg.llama.hidden_size, cfg.llama.intermediate_size
        if name.startswith(VIT):
            n = name[len(VIT):]
            simple = {"embeddings.class_embedding": "vit.cls", "embeddings.position_embedding.weight": "vit.pos",
                      "pre_layrnorm.weight": "vit.pre_ln_w", "pre_layrnorm.bias": "vit.pre_ln_b"}
            if n in simple:
                return t[simple[n]]
            if n == "embeddings.patch_embedding.weight":
                return None
            parts = n.split(".")                      # encoder.layers.i.<...>
            i, rest = int(parts[2]), ".".join(parts[3:])
            p = f"vit.{i}."
            qkv = {"q_proj": 0, "k_proj": 1, "v_proj": 2}
            if rest.startswith("self_attn.") and parts[4] in qkv:
                j = qkv[parts[4]]
                return t[p + ("qkv_w" if parts[5] == "weight" else "qkv_b")][j * d:(j + 1) * d]
            m = {"self_attn.out_proj.weight": "out_w", "self_attn.out_proj.bias": "out_b",
                 "layer_norm1.weight": "ln1_w", "layer_norm1.bias": "ln1_b", "layer_norm2.weight": "ln2_w",
                 "layer_norm2.bias": "ln2_b", "mlp.fc1.weight": "fc1_w", "mlp.fc1.bias": "fc1_b",
                 "mlp.fc2.weight": "fc2_w", "mlp.fc2.bias": "fc2_b"}
            return t[p + m[rest]]
        if name.startswith("model.mm_projector."):
            return t[{"0.weight": "proj.w0", "0.bias": "proj.b0", "2.weight": "proj.w2", "2.bias": "proj.b2"}[name[len("model.mm_projector."):]]]
        if name == "model.embed_tokens.weight":
            return t["llama.embed"]
        if name == "model.norm.weight":
            return t["llama.final_norm"]
        if name == "lm_head.weight":
            return t["llama.lm_head"]
        parts = name.split(".")                       # model.layers.i.<...>
        i, rest = int(parts[2]), ".".join(parts[3:])
        p = f"llama.{i}."
        qkv = {"self_attn.q_proj.weight": 0, "self_attn.k_proj.weight": 1, "self_attn.v_proj.weight": 2}
        if rest in qkv:
            j = qkv[rest]
            return t[p + "qkv_w"][j * h:(j + 1) * h]
        if rest == "mlp.gate_proj.weight":
            return t[p + "gate_up_w"][:I]
        if rest == "mlp.up_proj.weight":
            return t[p + "gate_up_w"][I:]
        return t[p + {"self_attn.o_proj.weight": "o_w", "mlp.down_proj.weight": "down_w",
                      "input_layernorm.weight": "in_norm", "post_attention_layernorm.weight": "post_norm"}[rest]]

    @classmethod
    def from_synthetic(cls, cfg: TeoConfig, seed: int, device) -> "TeoWeights":
        device = torch.device(device)
        self = cls(cfg, device)
        lib = L.load()
        stream = torch.cuda.current_stream(device).cuda_stream
        pdim = cfg.vision.patch_dim
        for name, shape, std, mean in param_specs(cfg):
            dst = self._dest(name)
            tmp = None
            if dst is None:        # patch weight: generate dense [d, 588], then place into [d, kpad]
                tmp = torch.empty(shape[0], pdim, dtype=torch.bfloat16, device=device)
                dst = tmp
            assert dst.is_contiguous() and dst.numel() == math.prod(shape), name
            L.check(lib.teo_init_normal_hash_bf16(dst.data_ptr(), dst.numel(), ctypes.c_uint64(tensor_seed(seed, name)),
                                                  hash_scale(std), float(mean), stream), f"init {name}")
            if tmp is not None:
                self.t["vit.patch_w"][:, :pdim] = tmp
        return self.to_blocked()

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], cfg: TeoConfig, device) -> "TeoWeights":
        device = torch.device(device)
        self = cls(cfg, device)
        pdim = cfg.vision.patch_dim
        for name, shape, _, _ in param_specs(cfg):
            if name not in sd:
                raise KeyError(f"state dict lacks {name}")
            src = sd[name]
            if tuple(src.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {shape}, got {tuple(src.shape)}")
            dst = self._dest(name)
            src = src.to(device=device, dtype=torch.bfloat16)
            if dst is None:
                self.t["vit.patch_w"][:, :pdim] = src.reshape(shape[0], pdim)
            else:
                dst.copy_(src.reshape(dst.shape))
        return self.to_blocked()

    # ---- blocked GEMM-weight layout (include/teochat_b200.h: teo_weight_to_blocked) -----------------------
    def _gemm_weight_groups(self):
        v, l = self.cfg.vision, self.cfg.llama
        folded = ("qkv_wf", "fc1_wf") if self.ln_folded else ()
        vit = ["vit.patch_w"] + [f"vit.{i}.{n}" for i in range(v.num_hidden_layers) for n in ("qkv_w", "out_w", "fc1_w", "fc2_w") + folded]
        proj = ["proj.w0", "proj.w2"]
        llama = [f"llama.{i}.{n}" for i in range(l.num_hidden_layers) for n in ("qkv_w", "o_w", "gate_up_w", "down_w")] + ["llama.lm_head"]
        return {"vit": vit, "proj": proj, "llama": llama}

    def interleave_gate_up(self) -> "TeoWeights":
        """gate_up_w [gate; up] → rows interleaved in blocks of 32 (| gate 32 | up 32 |…): accumulator columns c..c+31 and
        c+32..c+63 of the gate/up GEMM then belong to the same 32 outputs (TEO_ACT_SWIGLU_PAIRS).  Runs once, before the
        blocked re-layout; needs intermediate_size % 32 == 0 (else the layout stays [gate; up] and SwiGLU is a kernel)."""
        l = self.cfg.llama
        I = l.intermediate_size
        if self.gate_up_interleaved or self.blocked.get("llama") or I % 32 != 0:
            return self
        for i in range(l.num_hidden_layers):
            k = f"llama.{i}.gate_up_w"
            w = self.t[k]
            self.t[k] = w.view(2, I // 32, 32, w.shape[1]).permute(1, 0, 2, 3).contiguous().view(2 * I, w.shape[1])
        self.gate_up_interleaved = True
        return self

    def fold_vit_layernorm(self) -> "TeoWeights":
        """LayerNorm folded into the linear that consumes it (include/teochat_b200.h: teo_gemm_bf16_ex): per ViT layer
        qkv_wf = bf16(ln1_w ⊙ qkv_w), qkv_c = rowsum(qkv_wf) (f32), qkv_bf = qkv_w·ln1_b + qkv_b (f32), and fc1_* with ln2.
        Runs once, before the blocked re-layout; the plain weights stay (exact mode, small batches)."""
        if self.ln_folded or self.blocked.get("vit"):
            return self
        for i in range(self.cfg.vision.num_hidden_layers):
            p = f"vit.{i}."
            for lin, ln in (("qkv", "ln1"), ("fc1", "ln2")):
                w, b = self.t[p + lin + "_w"].float(), self.t[p + lin + "_b"].float()
                g, beta = self.t[p + ln + "_w"].float(), self.t[p + ln + "_b"].float()
                wf = (w * g[None, :]).to(torch.bfloat16)
                self.t[p + lin + "_wf"] = wf.contiguous()
                self.t[p + lin + "_c"] = wf.float().sum(dim=1).contiguous()
                self.t[p + lin + "_bf"] = (w @ beta + b).contiguous()
        self.ln_folded = True
        return self

    def to_blocked(self) -> "TeoWeights":
        """Re-lay every GEMM weight [N,K] as [N/128][K/64][128][64] (16 KiB contiguous operand tiles) when all
        matrices of a model part allow it; sets ``self.blocked[part]``.  Idempotent."""
        self.interleave_gate_up()
        if os.environ.get("TEO_VIT_LN_FOLD", "0") == "1":       # optional path, off by default (engine._build_structs)
            self.fold_vit_layernorm()
        for part, keys in self._gemm_weight_groups().items():
            if self.blocked.get(part):
                continue
            ok = all(self.t[k].shape[0] % 128 == 0 and self.t[k].shape[1] % 64 == 0 for k in keys)
            if ok:
                for k in keys:
                    w = self.t[k]
                    n, kk = w.shape
                    self.t[k] = w.view(n // 128, 128, kk // 64, 64).permute(0, 2, 1, 3).contiguous().view(n, kk)
            self.blocked[part] = ok
        return self

    def nbytes(self) -> int:
        return sum(x.numel() * x.element_size() for x in self.t.values())
