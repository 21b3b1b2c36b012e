"""Builds the in-tree CUDA shared library (sm_100a only) with nvcc.  No torch extension
machinery: the product boundary is a plain C-ABI .so loaded through ctypes."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libteochat_b200.so")
SOURCES = ["gemm.cu", "gemm_pair.cu", "attention.cu", "attention_tc.cu", "decode_attn_mma.cu", "kernels_misc.cu", "preprocess.cu", "exact.cu", "model_exact.cu", "model.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "teochat_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, defines) -> str:
    """Another build of the same sources with extra -D flags (compile-time A/B variants), as lib/variants/<name>.so;
    select it at run time with TEO_LIB_PATH."""
    vdir = os.path.join(LIB_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, f"{name}.so")
    cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], *[f"-D{d}" for d in defines], "-shared",
           *[os.path.join(CSRC, s) for s in SOURCES], "-o", out, "-cudart", "static"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed; see teochat_b200/lib/build.log")
    if verbose:
        print("\n".join(log))
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-cudart", "static"]
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
