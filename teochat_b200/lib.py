"""ctypes binding of the C-ABI in include/teochat_b200.h (the product's only compute path).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

OK = 0
ACT_NONE, ACT_QUICK_GELU, ACT_GELU, ACT_SWIGLU_PAIRS = 0, 1, 2, 3
ACT_BY_NAME = {"none": ACT_NONE, "quick_gelu": ACT_QUICK_GELU, "gelu": ACT_GELU}

vp = C.c_void_p


class TeoError(RuntimeError):
    pass


class GemmOpts(C.Structure):
    _fields_ = [("bias", vp), ("residual", vp), ("ldr", C.c_int), ("act", C.c_int), ("out_fp32", C.c_int), ("w_blocked", C.c_int),
                ("ln_stats", vp), ("ln_c", vp), ("ln_bias", vp), ("ln_slots", C.c_int), ("ln_eps", C.c_float), ("stats_out", vp)]


class VitLayer(C.Structure):
    _fields_ = [(n, vp) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b", "ln2_w", "ln2_b",
                                  "fc1_w", "fc1_b", "fc2_w", "fc2_b", "qkv_wf", "qkv_c", "qkv_bf", "fc1_wf", "fc1_c", "fc1_bf")]


class VitModel(C.Structure):
    _fields_ = [("hidden", C.c_int), ("inter", C.c_int), ("heads", C.c_int), ("image", C.c_int),
                ("patch", C.c_int), ("kpad", C.c_int), ("act", C.c_int), ("layers_run", C.c_int),
                ("eps", C.c_float), ("w_blocked", C.c_int), ("exact", C.c_int), ("patch_w", vp), ("cls", vp), ("pos", vp), ("pre_ln_w", vp),
                ("pre_ln_b", vp), ("layers", C.POINTER(VitLayer))]


class Projector(C.Structure):
    _fields_ = [("in_dim", C.c_int), ("hidden", C.c_int), ("w_blocked", C.c_int), ("exact", C.c_int), ("w0", vp), ("b0", vp), ("w2", vp), ("b2", vp)]


class LlamaLayer(C.Structure):
    _fields_ = [(n, vp) for n in ("in_norm", "qkv_w", "o_w", "post_norm", "gate_up_w", "down_w", "kv_pages")]


class LlamaModel(C.Structure):
    _fields_ = [("hidden", C.c_int), ("inter", C.c_int), ("heads", C.c_int), ("layers", C.c_int),
                ("vocab", C.c_int), ("page_size", C.c_int), ("rope_max_pos", C.c_int), ("eps", C.c_float),
                ("w_blocked", C.c_int), ("gate_up_interleaved", C.c_int), ("exact", C.c_int), ("rope_cos", vp), ("rope_sin", vp), ("embed", vp),
                ("final_norm", vp), ("lm_head", vp),
                ("layer", C.POINTER(LlamaLayer))]


i, f, sz, u64 = C.c_int, C.c_float, C.c_size_t, C.c_uint64
_SIGS = {
    # name: (restype, argtypes)
    "teo_create": (i, [i, C.POINTER(vp)]),
    "teo_destroy": (i, [vp]),
    "teo_last_error": (C.c_char_p, []),
    "teo_abi_version": (i, []),
    "teo_build_digest": (C.c_char_p, []),
    "teo_launch_count": (C.c_ulonglong, [vp]),
    "teo_init_normal_hash_bf16": (i, [vp, sz, u64, f, f, vp]),
    "teo_init_normal_hash_f32": (i, [vp, sz, u64, f, f, vp]),
    "teo_init_u8_hash": (i, [vp, sz, u64, vp]),
    "teo_gemm_workspace_bytes": (sz, [i, i, i]),
    "teo_gemm_bf16": (i, [vp, vp, i, vp, i, vp, i, i, i, i, vp, vp, i, i, i, vp, sz, vp]),
    "teo_gemm_stats_slots": (i, [i, i, i]),
    "teo_gemm_bf16_ex": (i, [vp, vp, i, vp, i, vp, i, i, i, i, C.POINTER(GemmOpts), vp, sz, vp]),
    "teo_row_stats": (i, [vp, vp, i, i, i, vp]),
    "teo_gemm_bf16_wblocked": (i, [vp, vp, i, vp, vp, i, i, i, i, vp, vp, i, i, i, vp, sz, vp]),
    "teo_weight_to_blocked": (i, [vp, vp, i, i, vp]),
    "teo_patchify_u8_nhwc": (i, [vp, vp, i, i, i, i, vp]),
    "teo_patchify_f32_nchw": (i, [vp, vp, i, i, i, i, vp]),
    "teo_resize_workspace_bytes": (sz, [i, i, i]),
    "teo_resize_crop_normalize_u8": (i, [vp, i, i, i, i, i, i, i, C.POINTER(f), C.POINTER(f), vp, vp, sz, vp]),
    "teo_vit_assemble_preln": (i, [vp, vp, vp, vp, vp, vp, i, i, i, f, vp]),
    "teo_layernorm": (i, [vp, vp, vp, vp, i, i, f, vp]),
    "teo_vit_drop_cls": (i, [vp, vp, i, i, i, vp]),
    "teo_flash_attention": (i, [vp, i, vp, i, vp, i, vp, i, vp, i, i, i, i, f, i, vp]),
    "teo_flash_attention_tc": (i, [vp, vp, i, vp, i, vp, i, vp, i, vp, i, i, i, i, i, f, i, i, vp]),
    "teo_kv_pool_bytes": (sz, [i, i, i, i, i]),
    "teo_kv_plan": (i, [C.POINTER(i), i, i, i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
    "teo_kv_create": (i, [i, C.POINTER(vp)]),
    "teo_kv_destroy": (i, [vp]),
    "teo_kv_available": (i, [vp]),
    "teo_kv_alloc": (i, [vp, i, i, C.POINTER(i), i]),
    "teo_kv_free": (i, [vp, C.POINTER(i), i]),
    "teo_rope_kv_write": (i, [vp, vp, vp, vp, vp, i, i, i, i, i, vp, vp, vp]),
    "teo_decode_attention_workspace_bytes": (sz, [i, i, i, i]),
    "teo_decode_attention": (i, [vp, i, vp, vp, i, vp, vp, i, i, i, i, i, f, vp, sz, vp]),
    "teo_decode_attention_h": (i, [vp, vp, i, vp, vp, i, vp, vp, i, i, i, i, i, f, vp, sz, vp]),
    "teo_rmsnorm": (i, [vp, vp, vp, i, i, f, vp]),
    "teo_swiglu": (i, [vp, vp, i, i, vp]),
    "teo_splice_embed": (i, [vp, vp, vp, vp, i, i, vp]),
    "teo_splice_embed_f32": (i, [vp, vp, vp, vp, i, i, vp]),
    "teo_split_f32_bf16x3": (i, [vp, vp, i, i, vp]),
    "teo_gemm_bf16x3": (i, [vp, vp, vp, i, vp, i, i, i, vp, vp, vp, sz, vp]),
    "teo_argmax_step": (i, [vp, i, vp, vp, i, i, vp, i, i, vp]),
    "teo_sample_step": (i, [vp, i, f, i, u64, vp, vp, i, i, vp, i, i, vp]),
    "teo_set_sampling": (i, [vp, f, i, u64]),
    "teo_set_sampling_seed_device": (i, [vp, vp]),
    "teo_set_decode_chain": (i, [vp, i]),
    "teo_set_pdl": (i, [vp, i]),
    "teo_vit_workspace_bytes": (sz, [C.POINTER(VitModel), i]),
    "teo_vit_encode": (i, [vp, C.POINTER(VitModel), vp, vp, i, vp, vp, sz, vp]),
    "teo_projector_workspace_bytes": (sz, [C.POINTER(Projector), i]),
    "teo_projector_mlp2x": (i, [vp, C.POINTER(Projector), vp, i, vp, vp, sz, vp]),
    "teo_llama_prefill_workspace_bytes": (sz, [C.POINTER(LlamaModel), i, i]),
    "teo_llama_prefill": (i, [vp, C.POINTER(LlamaModel), vp, i, vp, vp, vp, vp, i, i, vp, i, vp, vp, sz, vp]),
    "teo_llama_decode_workspace_bytes": (sz, [C.POINTER(LlamaModel), i, i]),
    "teo_llama_decode_step": (i, [vp, C.POINTER(LlamaModel), vp, vp, vp, vp, i, vp, i, i, vp, i, vp, i, vp, sz, vp]),
}
EXPORTS = tuple(_SIGS)


def lib_path() -> str:
    """The in-tree library; TEO_LIB_PATH selects another build of the same sources (A/B measurements of compile-time
    variants, scripts/gpu_dec_ab.sh)."""
    return os.environ.get("TEO_LIB_PATH") or _build.LIB_PATH


def load() -> C.CDLL:
    """Load (never build implicitly on a GPU box: the .so travels with the snapshot)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise TeoError(f"{path} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                           "There is no CPU or PyTorch fallback.")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        if not os.environ.get("TEO_LIB_PATH"):       # the in-tree build must come from exactly the sources beside it
            have, want = lib.teo_build_digest().decode(), _build.source_digest()
            if have != want:
                raise TeoError(f"{path} was built from other sources (digest {have}, sources {want}): "
                               "run `python __graft_entry__.py build`")
        _LIB = lib
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        msg = load().teo_last_error().decode("utf-8", "replace")
        raise TeoError(f"{what or 'teochat_b200 call'} failed ({rc}): {msg}")


def ptr(t) -> int:
    """device (or host) address of a torch tensor, None → NULL"""
    return None if t is None else t.data_ptr()
