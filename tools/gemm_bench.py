"""Times the tiled GEMM (CTA-pair kernel; TEO_GEMM_PAIR=0 → single-CTA) at the ViT / prefill shapes with different epilogues, to
tell mainloop-bound from epilogue-bound.  Development tool: `python tools/gemm_bench.py [vit|prefill]` on a B200."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teochat_b200 import lib as L  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "vit"
    lib = L.load()
    h = C.c_void_p()
    L.check(lib.teo_create(0, C.byref(h)))
    st = torch.cuda.current_stream().cuda_stream
    dev = "cuda"
    if which == "vit":
        M, shapes = 256 * 257, [("qkv", 3072, 1024), ("out", 1024, 1024), ("fc1", 4096, 1024), ("fc2", 1024, 4096)]
    else:
        M, shapes = 68160, [("qkv", 12288, 4096), ("o", 4096, 4096), ("gate_up", 22016, 4096), ("down", 4096, 11008)]
    for name, N, K in shapes:
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        W = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5
        Wb = torch.empty_like(W)
        L.check(lib.teo_weight_to_blocked(W.data_ptr(), Wb.data_ptr(), N, K, st))
        bias = torch.randn(N, device=dev, dtype=torch.bfloat16)
        res = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        row = []
        for label, b, r, act in (("plain", None, None, 0), ("bias", bias, None, 0), ("bias+quick_gelu", bias, None, 1), ("bias+residual", bias, res, 0)):
            def run():
                L.check(lib.teo_gemm_bf16_wblocked(h, A.data_ptr(), K, Wb.data_ptr(), out.data_ptr(), N, M, N, K, L.ptr(b), L.ptr(r), N, act, 0,
                                                   None, 0, st))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            it = 10
            e0.record()
            for _ in range(it):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / it * 1e3
            row.append(f"{label} {us:7.1f} us {2.0 * M * N * K / us / 1e6:6.0f} TF/s")
        print(f"{which} {name:8s} M={M} N={N} K={K}: " + " | ".join(row), flush=True)
        del A, W, Wb, res, out


if __name__ == "__main__":
    main()
