"""Sweep of the CTA-pair GEMM's rasterisation (supertile group_m × group_n, band / strip order) and L2 eviction hints at the
prefill shapes of the bench step (M = 68160 tokens), through the development hook teo_dbg_pair_cfg.  Development tool:
    python tools/pair_sweep.py time              CUDA-event timing of every configuration
    python tools/pair_sweep.py one gm gn r h     one launch per shape of one configuration (for an ncu metrics pass)
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teochat_b200 import lib as L  # noqa: E402

M = 68160
SHAPES = [("gate_up", 22016, 4096, 3), ("down", 4096, 11008, 0), ("qkv", 12288, 4096, 0), ("o", 4096, 4096, 0)]
CONFIGS = [(16, 1 << 20, 0, 0), (16, 1 << 20, 0, 1), (16, 16, 0, 1), (24, 24, 0, 0), (24, 24, 0, 1), (32, 16, 0, 1), (32, 1 << 20, 0, 1),
           (16, 16, 1, 0), (16, 16, 1, 1), (24, 16, 1, 1), (8, 32, 1, 1), (48, 1 << 20, 0, 1)]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "time"
    lib = L.load()
    raw = C.CDLL(L.lib_path())
    raw.teo_dbg_pair_cfg.argtypes = [C.c_int] * 4
    h = C.c_void_p()
    L.check(lib.teo_create(0, C.byref(h)))
    st = torch.cuda.current_stream().cuda_stream
    dev = "cuda"
    cfgs = CONFIGS if mode == "time" else [tuple(int(x) for x in sys.argv[2:6])]
    for name, N, K, act in SHAPES:
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        W = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5
        Wb = torch.empty_like(W)
        L.check(lib.teo_weight_to_blocked(W.data_ptr(), Wb.data_ptr(), N, K, st))
        n_out = N // 2 if act == 3 else N
        out = torch.empty(M, n_out, device=dev, dtype=torch.bfloat16)
        res = out if name in ("down", "o") else None

        def run():
            L.check(lib.teo_gemm_bf16_wblocked(h, A.data_ptr(), K, Wb.data_ptr(), out.data_ptr(), n_out, M, N, K, None, L.ptr(res), n_out if res is not None else 0,
                                               act, 0, None, 0, st))
        for gm, gn, raster, hint in cfgs:
            raw.teo_dbg_pair_cfg(gm, gn, raster, hint)
            if mode != "time":
                torch.cuda.profiler.start()
                run()
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()
                continue
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            it = 6
            e0.record()
            for _ in range(it):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / it * 1e3
            print(f"{name:8s} N={N:5d} K={K:5d} gm={gm:2d} gn={gn if gn < 1000 else 'all':>3} raster={raster} hint={hint}: {us:8.1f} us "
                  f"{2.0 * M * N * K / us / 1e6:6.0f} TF/s", flush=True)
        del A, W, Wb, out


if __name__ == "__main__":
    main()
