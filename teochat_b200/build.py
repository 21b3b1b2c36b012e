"""Builds the in-tree CUDA shared library (sm_100a only) with nvcc.  No torch extension
machinery: the product boundary is a plain C-ABI .so loaded through ctypes."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libteochat_b200.so")
SOURCES = ["gemm.cu", "gemm_pair.cu", "attention.cu", "attention_tc.cu", "decode_attn_mma.cu", "decode_chain.cu", "kernels_misc.cu", "preprocess.cu", "exact.cu", "model_exact.cu", "kv_pages.cu", "model.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


DIGEST_PATH = os.path.join(LIB_DIR, "libteochat_b200.digest")


def source_digest() -> str:
    """sha256 over everything the library is built from: csrc/* (names + bytes), the public header and the nvcc flags.
    The digest is compiled INTO the library (teo_build_digest()) and lib.load() refuses a library whose digest differs from
    the sources beside it — a stale .so cannot be picked up silently, wherever it was built."""
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(HERE, "..", "include", "teochat_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()[:16]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(DIGEST_PATH):
        return True
    with open(DIGEST_PATH) as f:
        return f.read().strip() != source_digest()


def build_variant(name: str, defines) -> str:
    """Another build of the same sources with extra -D flags (compile-time A/B variants), as lib/variants/<name>.so;
    select it at run time with TEO_LIB_PATH."""
    vdir = os.path.join(LIB_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, f"{name}.so")
    cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], *[f"-D{d}" for d in defines],
           f'-DTEO_BUILD_DIGEST="{source_digest()}"', "-shared",
           *[os.path.join(CSRC, s) for s in SOURCES], "-o", out, "-cudart", "static"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = source_digest()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, f'-DTEO_BUILD_DIGEST="{digest}"', "-c", os.path.join(CSRC, src), "-o", obj]
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
    with open(DIGEST_PATH, "w") as f:
        f.write(digest + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
