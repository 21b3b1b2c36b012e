"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/teochat_b200.h declares, and fails loudly (no fallback) when asked to compute without one."""
import ctypes as C
import os
import re

import pytest
import torch

from teochat_b200 import build as B
from teochat_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(B.LIB_PATH):
        B.build()
    return L.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "teochat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(teo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/teochat_b200.h but not exported"
    assert set(syms) == set(L.EXPORTS), "ctypes signature table and header disagree"
    assert lib.teo_abi_version() == 4


def test_struct_layouts_match_header():
    # pointer-sized fields, int fields in header order (catches drift between lib.py and the header)
    assert C.sizeof(L.VitLayer) == 18 * C.sizeof(C.c_void_p)          # 12 parameters + the 6 folded-LayerNorm tensors
    assert C.sizeof(L.LlamaLayer) == 7 * C.sizeof(C.c_void_p)
    assert [f for f, _ in L.VitModel._fields_][:11] == ["hidden", "inter", "heads", "image", "patch", "kpad", "act", "layers_run", "eps",
                                                        "w_blocked", "exact"]
    assert [f for f, _ in L.Projector._fields_][:4] == ["in_dim", "hidden", "w_blocked", "exact"]
    assert [f for f, _ in L.LlamaModel._fields_][:13] == ["hidden", "inter", "heads", "layers", "vocab", "page_size", "rope_max_pos", "eps",
                                                         "w_blocked", "gate_up_interleaved", "exact", "rope_cos", "rope_sin"]
    # field order in the header text itself
    hdr = open(os.path.join(ROOT, "include", "teochat_b200.h")).read()
    for struct, fields in (("teo_vit_model", ["eps;", "w_blocked;", "exact;", "patch_w;"]), ("teo_projector", ["w_blocked;", "exact;", "*w0"]),
                           ("teo_llama_model", ["eps;", "w_blocked;", "gate_up_interleaved;", "exact;", "*rope_cos"])):
        body = hdr[hdr.rindex("typedef struct {", 0, hdr.index("} " + struct + ";")):hdr.index("} " + struct + ";")]
        pos = [body.index(f) for f in fields]
        assert pos == sorted(pos), struct


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(lib):
    h = C.c_void_p()
    rc = lib.teo_create(0, C.byref(h))
    assert rc != 0 and len(lib.teo_last_error()) > 0
    from teochat_b200.config import TeoConfig
    from teochat_b200.engine import TeoModel
    with pytest.raises(L.TeoError, match="no CPU path"):
        TeoModel(TeoConfig.tiny(), None, "cuda:0")
    from teochat_b200.eval.eval import load_model
    with pytest.raises(ValueError, match="llava"):
        load_model("some-other-model", None)
    with pytest.raises(NotImplementedError):
        load_model("teochat-synthetic", None, load_8bit=True)


def test_workspace_queries_are_pure(lib):
    assert lib.teo_gemm_workspace_bytes(4096, 4096, 4096) == 0            # large M: no split-K scratch
    assert lib.teo_gemm_workspace_bytes(32, 4096, 4096) == 19 * 32 * 4096 * 4     # partial slots of the small-M schedule
    assert lib.teo_decode_attention_workspace_bytes(32, 32, 128, 8) == 32 * 32 * 8 * 130 * 4


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "teochat_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_kv_page_planner_and_allocator(lib):
    """teo_kv_plan / teo_kv_create / teo_kv_alloc / teo_kv_free (SURVEY.md §8b minimum set): host-only, so it runs without a GPU."""
    lens = (C.c_int * 4)(130, 64, 1, 2130)
    per, mx, tot = (C.c_int * 4)(), C.c_int(), C.c_int()
    assert lib.teo_kv_plan(lens, 4, 256, 64, per, C.byref(mx), C.byref(tot)) == 0
    assert list(per) == [7, 5, 5, 38] and mx.value == 38 and tot.value == 55         # ceil((len + 256) / 64)
    assert lib.teo_kv_plan(lens, 4, 256, 0, per, None, None) != 0 and b"kv_plan" in lib.teo_last_error()
    assert lib.teo_kv_pool_bytes(55, 32, 64, 128, 0) == 55 * 2 * 32 * 64 * 128 * 2
    assert lib.teo_kv_pool_bytes(55, 32, 64, 128, 1) == 55 * 2 * 32 * 64 * 128 * 4
    a = C.c_void_p()
    assert lib.teo_kv_create(12, C.byref(a)) == 0 and lib.teo_kv_available(a) == 12
    row = (C.c_int * 8)()
    assert lib.teo_kv_alloc(a, 130 + 256, 64, row, 8) == 7 and list(row)[:7] == [0, 1, 2, 3, 4, 5, 6]      # lowest ids first
    row2 = (C.c_int * 8)()
    assert lib.teo_kv_alloc(a, 200, 64, row2, 8) == 4 and list(row2)[:4] == [7, 8, 9, 10]
    assert lib.teo_kv_alloc(a, 129, 64, row2, 8) == -3 and lib.teo_kv_available(a) == 1         # exhausted: nothing taken
    assert lib.teo_kv_alloc(a, 64 * 9, 64, row2, 8) == -1                                        # row too short for 9 pages
    assert lib.teo_kv_free(a, row, 7) == 0 and lib.teo_kv_available(a) == 8
    assert lib.teo_kv_free(a, row, 7) == -1                                                      # double free is rejected
    row3 = (C.c_int * 8)()
    assert lib.teo_kv_alloc(a, 3 * 64, 64, row3, 8) == 3 and list(row3)[:3] == [0, 1, 2]          # freed pages are reused
    assert lib.teo_kv_destroy(a) == 0
