"""Sweep of the CTA-pair GEMM's rasterisation (supertile group_m × group_n, band / strip order) and L2 eviction hints at the
prefill / ViT shapes of the bench step, through the development hook teo_dbg_pair_cfg.  Configurations are timed ROUND-ROBIN
(every round visits every configuration, medians reported) so that clock / power drift of the card does not favour whoever
runs first.  Development tool:
    python tools/pair_sweep.py time [prefill|vit]      CUDA-event timing of every configuration
    python tools/pair_sweep.py one gm gn r h [which]    one launch per shape of one configuration (for an ncu metrics pass)
    python tools/pair_sweep.py cublas [prefill|vit]    this kernel against torch.matmul (cuBLASLt) on the same operands, interleaved
"""
import ctypes as C
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teochat_b200 import lib as L  # noqa: E402

ALL = 1 << 20
SHAPES = {
    "prefill": (68160, [("gate_up", 22016, 4096, 3), ("down", 4096, 11008, 0), ("qkv", 12288, 4096, 0), ("o", 4096, 4096, 0)]),
    "vit": (256 * 257, [("qkv", 3072, 1024, 0), ("out", 1024, 1024, 0), ("fc1", 4096, 1024, 1), ("fc2", 1024, 4096, 0)]),
}
CONFIGS = [(16, ALL, 0, 0), (8, ALL, 0, 0), (12, ALL, 0, 0), (24, ALL, 0, 0), (32, ALL, 0, 0), (16, ALL, 0, 2), (24, ALL, 0, 2), (32, ALL, 0, 2),
           (16, 16, 1, 0), (16, 16, 1, 2), (16, 8, 1, 2), (24, 24, 0, 0), (16, ALL, 0, 3)]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "time"
    which = (sys.argv[2] if mode in ("time", "cublas") and len(sys.argv) > 2 else (sys.argv[6] if mode == "one" and len(sys.argv) > 6 else "prefill"))
    lib = L.load()
    raw = C.CDLL(L.lib_path())
    raw.teo_dbg_pair_cfg.argtypes = [C.c_int] * 4
    h = C.c_void_p()
    L.check(lib.teo_create(0, C.byref(h)))
    st = torch.cuda.current_stream().cuda_stream
    dev = "cuda"
    cfgs = [tuple(int(x) for x in sys.argv[2:6])] if mode == "one" else CONFIGS
    M, shapes = SHAPES[which]
    for name, N, K, act in shapes:
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        W = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5
        Wb = torch.empty_like(W)
        L.check(lib.teo_weight_to_blocked(W.data_ptr(), Wb.data_ptr(), N, K, st))
        n_out = N // 2 if act == 3 else N
        out = torch.empty(M, n_out, device=dev, dtype=torch.bfloat16)
        res = out if name in ("down", "o", "out", "fc2") else None

        def run():
            L.check(lib.teo_gemm_bf16_wblocked(h, A.data_ptr(), K, Wb.data_ptr(), out.data_ptr(), n_out, M, N, K, None, L.ptr(res),
                                               n_out if res is not None else 0, act, 0, None, 0, st))
        if mode == "cublas":
            # the library's rate on the same operands, interleaved with ours (same clocks / power state): torch.matmul → cuBLASLt
            Wt = W.t()
            raw.teo_dbg_pair_cfg(0, 0, -1, -1)
            ours, lib_t = [], []
            for _ in range(6):
                run()
                torch.matmul(A, Wt)
            torch.cuda.synchronize()
            for rnd in range(6):
                for which_one in ((0, 1) if rnd % 2 == 0 else (1, 0)):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        if which_one == 0:
                            run()
                        else:
                            torch.matmul(A, Wt)
                    e1.record()
                    torch.cuda.synchronize()
                    (ours if which_one == 0 else lib_t).append(e0.elapsed_time(e1) / 3 * 1e3)
            a_us, b_us = statistics.median(ours), statistics.median(lib_t)
            fl = 2.0 * M * N * K
            print(f"{which} {name:8s} M={M} N={N:5d} K={K:5d}: gemm_pair_kernel (+ its epilogue: {'SwiGLU' if act == 3 else ('residual' if res is not None else ('quick_gelu' if act == 1 else 'plain'))}) "
                  f"{a_us:8.1f} us {fl / a_us / 1e6:6.0f} TF/s | cuBLAS plain GEMM {b_us:8.1f} us {fl / b_us / 1e6:6.0f} TF/s | ours/cuBLAS time {a_us / b_us:5.3f}", flush=True)
            del A, W, Wb, out
            continue
        if mode != "time":
            raw.teo_dbg_pair_cfg(*cfgs[0])
            torch.cuda.profiler.start()
            run()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            continue
        for _ in range(6):              # bring the card to its sustained (power-capped) state first
            run()
        torch.cuda.synchronize()
        times = {c: [] for c in cfgs}
        for rnd in range(5):
            for c in (cfgs if rnd % 2 == 0 else cfgs[::-1]):
                raw.teo_dbg_pair_cfg(*c)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    run()
                e1.record()
                torch.cuda.synchronize()
                times[c].append(e0.elapsed_time(e1) / 3 * 1e3)
        base = statistics.median(times[cfgs[0]])
        for c in cfgs:
            us = statistics.median(times[c])
            gm, gn, raster, hint = c
            print(f"{which} {name:8s} N={N:5d} K={K:5d} gm={gm:2d} gn={gn if gn < 1000 else 'all':>3} raster={raster} hint={hint}: {us:8.1f} us "
                  f"{2.0 * M * N * K / us / 1e6:6.0f} TF/s  {100.0 * (us / base - 1):+5.1f}%  (min {min(times[c]):.0f} max {max(times[c]):.0f})", flush=True)
        del A, W, Wb, out
    raw.teo_dbg_pair_cfg(0, 0, -1, -1)


if __name__ == "__main__":
    main()
