"""Per-tile timeline of gemm_pair_kernel from clock64 stamps (development build: build.build_variant("pairtrace", ["TEO_PAIR_TRACE"]),
run with TEO_LIB_PATH=teochat_b200/lib/variants/pairtrace.so; every stamp costs ≈ 300 cycles itself).  `python tools/pair_trace.py [prefill|vit] [shape name]` on a B200."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teochat_b200 import lib as L  # noqa: E402
from pair_sweep import SHAPES  # noqa: E402

NAMES = ["issuer past tempty", "issuer has kb0", "issuer committed tfull", "epilogue saw tfull", "epilogue released", "epilogue done",
         "producer issued kb0", "producer issued last kb", "slot of kb0 free again", "slot of kb STAGES free again"]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "prefill"
    only = sys.argv[2] if len(sys.argv) > 2 else None
    lib = L.load()
    raw = C.CDLL(L.lib_path())
    raw.teo_dbg_pair_trace.argtypes = [C.c_void_p, C.c_int]
    h = C.c_void_p()
    L.check(lib.teo_create(0, C.byref(h)))
    st = torch.cuda.current_stream().cuda_stream
    M, shapes = SHAPES[which]
    T = 48
    for name, N, K, act in shapes:
        if only and name != only:
            continue
        A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
        W = torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * K ** -0.5
        Wb = torch.empty_like(W)
        L.check(lib.teo_weight_to_blocked(W.data_ptr(), Wb.data_ptr(), N, K, st))
        n_out = N // 2 if act == 3 else N
        out = torch.empty(M, n_out, device="cuda", dtype=torch.bfloat16)
        res = out if name in ("down", "o", "out", "fc2") else None

        def run():
            L.check(lib.teo_gemm_bf16_wblocked(h, A.data_ptr(), K, Wb.data_ptr(), out.data_ptr(), n_out, M, N, K, None, L.ptr(res),
                                               n_out if res is not None else 0, act, 0, None, 0, st))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        buf = torch.zeros(148 * T * 16, dtype=torch.int64, device="cuda")
        assert raw.teo_dbg_pair_trace(buf.data_ptr(), T) == 0
        run()
        torch.cuda.synchronize()
        raw.teo_dbg_pair_trace(None, 0)
        tr = buf.view(148, T, 16).cpu()
        print(f"=== {which} {name} M={M} N={N} K={K}: k-blocks {K // 64}")
        for cta in (0, 74):
            t = tr[cta]
            n = int((t[:, 0] > 0).sum())
            print(f"  CTA {cta} (leader of pair {cta // 2}): {n} traced tiles; cycles relative to the previous tile's tfull commit")
            for i in range(1, min(n, 9)):
                base = int(t[i - 1, 2])
                row = {k: int(t[i, k]) - base for k in range(10)}
                prev = {k: int(t[i - 1, k]) - base for k in (3, 4, 5)}
                e = {k: int(t[i - 1, k]) - base for k in range(10, 16)}
                print(f"    tile {i}: prev epilogue chunk 0: start {e[10]:+6d}, buffer free {e[11]:+6d}, TMEM read done {e[12]:+6d}, residual there {e[13]:+6d}, converted+staged {e[14]:+6d}, fenced {e[15]:+6d}")
                print(f"    tile {i}: prev epilogue saw tfull {prev[3]:+6d}, released {prev[4]:+6d}, done {prev[5]:+6d} | issuer past tempty {row[0]:+6d}, has kb0 {row[1]:+6d}, "
                      f"committed tfull {row[2]:+7d} (tile period) | producer kb0 issued {row[6]:+6d}, slot of kb0 free {row[8]:+6d}, slot of kb3 free {row[9]:+6d}, last kb issued {row[7]:+7d}")
            if n > 4:
                per = [(int(t[i, 2]) - int(t[i - 1, 2])) for i in range(2, n)]
                gap = [(int(t[i, 0]) - int(t[i - 1, 2])) for i in range(2, n)]
                print(f"    tile period: median {sorted(per)[len(per) // 2]} cycles (min {min(per)}, max {max(per)}); tfull commit -> next tile's first MMA issue: median {sorted(gap)[len(gap) // 2]}")
        del A, W, Wb, out


if __name__ == "__main__":
    main()
