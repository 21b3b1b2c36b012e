"""Times the four decode GEMMs of a LLaMA-2-7B layer alone (M = batch rows, swap-AB + stream-K + split-K reduce) and prints
the in-kernel timeline (%globaltimer stamps through the development hook teo_dbg_gemm_trace): when CTAs enter, when the
ring is filled, when the first operands land, last MMA, epilogue end, exit — relative to the first CTA's entry.
Weights rotate over enough copies to exceed the 126 MB L2.  Development tool: `python tools/dec_gemm_bench.py [M]`."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teochat_b200 import lib as L  # noqa: E402


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    lib = L.load()
    h = C.c_void_p()
    L.check(lib.teo_create(0, C.byref(h)))
    lib.teo_dbg_gemm_trace.restype = C.c_longlong
    lib.teo_dbg_gemm_trace.argtypes = [C.c_void_p, C.c_int]
    st = torch.cuda.current_stream().cuda_stream
    dev = "cuda"
    for name, N, K in (("qkv", 12288, 4096), ("o", 4096, 4096), ("gate_up", 22016, 4096), ("down", 4096, 11008)):
        copies = max(2, int(400e6 // (N * K * 2)) + 1)
        W = torch.randn(copies, N, K, device=dev, dtype=torch.bfloat16)
        Wb = torch.empty_like(W)
        for c in range(copies):
            L.check(lib.teo_weight_to_blocked(W[c].data_ptr(), Wb[c].data_ptr(), N, K, st))
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ws = torch.empty(lib.teo_gemm_workspace_bytes(M, N, K), dtype=torch.uint8, device=dev)

        def run(i):
            L.check(lib.teo_gemm_bf16_wblocked(h, A.data_ptr(), K, Wb[i % copies].data_ptr(), out.data_ptr(), N, M, N, K, None, None, 0, 0, 0,
                                               ws.data_ptr(), ws.numel(), st))
        for i in range(copies):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 40
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        ideal = N * K * 2 / 6.45e12 * 1e6
        trace = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
        lib.teo_dbg_gemm_trace(trace.data_ptr(), 1)
        run(1)
        torch.cuda.synchronize()
        lib.teo_dbg_gemm_trace(None, 0)
        t = trace.view(148, 8).cpu()
        t = t[t[:, 0] > 0]
        t0 = int(t[:, 0].min())
        rel = (t - t0).float() / 1e3
        names = ["enter", "(smid)", "ring filled", "dep wait done", "first operands", "last MMA issued", "epilogue done", "exit"]
        print(f"{name:8s} M={M} N={N} K={K}: GEMM + reduce {us:6.1f} us per call ({ideal:5.1f} us of weight streaming at 6.45 TB/s); "
              f"timeline over {t.shape[0]} CTAs, us from first entry (min / median / max):")
        for j, n in enumerate(names):
            col = rel[:, j]
            print(f"    {n:16s} {col.min():7.2f} {col.median():7.2f} {col.max():7.2f}")
        del W, Wb


if __name__ == "__main__":
    main()
