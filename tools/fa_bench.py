"""Times the two flash-attention kernels (mma.sync vs tcgen05) alone at the bench shapes.
Development tool: `python tools/fa_bench.py [vit|prefill|all]` on a B200."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teochat_b200 import lib as L  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    lib = L.load()
    h = C.c_void_p()
    L.check(lib.teo_create(0, C.byref(h)))
    st = torch.cuda.current_stream().cuda_stream
    shapes = []
    if which in ("vit", "all"):
        shapes.append(("vit", 64, 16, [257] * 256, False, 1))
        shapes.append(("vit_q0", 64, 16, [257] * 256, False, 0))       # everything tiled: third tile holds one row
        shapes.append(("vit256", 64, 16, [256] * 256, False, 0))       # no CLS: two full tiles, two full key blocks
    if which in ("prefill", "all"):
        shapes.append(("prefill", 128, 32, [2130] * 32, True, 0))
    for name, hd, H, lens, causal, qoff in shapes:
        T, d = sum(lens), H * hd
        qkv = torch.randn(T, 3 * d, device="cuda", dtype=torch.bfloat16)
        cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
        out = torch.empty(T, d, dtype=torch.bfloat16, device="cuda")
        out2 = torch.empty_like(out)
        args = (qkv.data_ptr(), 3 * d, qkv[:, d:].data_ptr(), 3 * d, qkv[:, 2 * d:].data_ptr(), 3 * d)
        flops = sum(4.0 * n * n * hd * H for n in lens) * (0.5 if causal else 1.0)

        def run_mma():
            L.check(lib.teo_flash_attention(*args, out.data_ptr(), d, cu.data_ptr(), len(lens), max(lens), H, hd, hd ** -0.5, int(causal), st))

        def run_tc():
            L.check(lib.teo_flash_attention_tc(h, *args, out2.data_ptr(), d, cu.data_ptr(), len(lens), max(lens), T, H, hd, hd ** -0.5,
                                               int(causal), qoff, st))

        for label, fn in (("mma.sync", run_mma), ("tcgen05", run_tc)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            print(f"{name:8s} {label:9s} {ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)
        diff = (out.float() - out2.float()).abs().max().item()
        print(f"{name:8s} max |mma - tc| = {diff:.4g} (max |out| {out.float().abs().max().item():.3g})", flush=True)


if __name__ == "__main__":
    main()
