"""Is the finish-time skew of the decode (stream-K, weight-streaming) GEMMs systematic per SM?  Runs each of the four decode GEMM
shapes many times with the in-kernel trace (teo_dbg_gemm_trace: %globaltimer stamps + %smid per CTA) and reports, per shape, the
CTAs' streaming time (dependency resolved → epilogue done) — mean / slowest — and how much of its variation is explained by
WHICH SM the CTA ran on (stable across launches) rather than by the launch.  `python tools/dec_gemm_skew.py [M]` on a B200."""
import ctypes as C
import os
import sys

import numpy as np
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
    n_launch = 48
    per_sm_all = []
    for name, N, K in (("qkv", 12288, 4096), ("o", 4096, 4096), ("gate_up", 22016, 4096), ("down", 4096, 11008)):
        copies = max(2, int(400e6 // (N * K * 2)) + 1)
        W = torch.randn(copies, N, K, device=dev, dtype=torch.bfloat16)
        Wb = torch.empty_like(W)
        for c in range(copies):
            L.check(lib.teo_weight_to_blocked(W[c].data_ptr(), Wb[c].data_ptr(), N, K, st))
        del W
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ws = torch.empty(lib.teo_gemm_workspace_bytes(M, N, K), dtype=torch.uint8, device=dev)

        def run(i):
            L.check(lib.teo_gemm_bf16_wblocked(h, A.data_ptr(), K, Wb[i % copies].data_ptr(), out.data_ptr(), N, M, N, K, None, None, 0, 0, 0,
                                               ws.data_ptr(), ws.numel(), st))
        for i in range(copies + 4):
            run(i)
        torch.cuda.synchronize()
        trace = torch.zeros(2 * n_launch, 148, 8, dtype=torch.int64, device=dev)     # every call = GEMM launch (+ a reduce kernel, not traced)
        lib.teo_dbg_gemm_trace(trace.data_ptr(), 2 * n_launch)
        for i in range(n_launch):
            run(i)
        torch.cuda.synchronize()
        n = lib.teo_dbg_gemm_trace(None, 0)
        t = trace.cpu().numpy()[:n].astype(np.float64)
        ok = t[:, :, 0] > 0
        ctas = int(ok[0].sum())
        t = t[:, :ctas]
        smid = t[:, :, 1].astype(int)
        start = t[:, :, 3]                                  # dependency (previous kernel) resolved
        dur = (t[:, :, 6] - np.minimum(start, t[:, :, 4])) / 1e3          # → this CTA's epilogue done, us
        span = (t[:, :, 7].max(1) - t[:, :, 0].min(1)) / 1e3
        stable = (smid == smid[0]).all()
        # per-SM mean over launches (CTA index ↔ SM may change between launches: group by smid)
        sm_ids = np.unique(smid)
        per_sm = np.array([dur[smid == s].mean() for s in sm_ids])
        resid = dur - np.array([per_sm[np.searchsorted(sm_ids, s)] for s in smid.ravel()]).reshape(dur.shape)
        work = np.ones(ctas)
        print(f"{name:8s} M={M} N={N} K={K}: {n} launches x {ctas} CTAs; kernel span {np.median(span):6.1f} us; CTA stream time mean {dur.mean():6.2f} "
              f"slowest-per-launch {np.median(dur.max(1)):6.2f} fastest {np.median(dur.min(1)):6.2f} us; blockIdx->SM mapping stable: {stable}")
        print(f"          std of CTA time: total {dur.std():5.2f} us, explained by SM {per_sm.std():5.2f} us, residual (per launch) {resid.std():5.2f} us; "
              f"per-SM mean: min {per_sm.min():6.2f} (sm {sm_ids[per_sm.argmin()]}) max {per_sm.max():6.2f} (sm {sm_ids[per_sm.argmax()]})")
        order = np.argsort(per_sm)
        print("          slowest SMs:", [(int(sm_ids[i]), round(float(per_sm[i]), 1)) for i in order[-12:]])
        print("          fastest SMs:", [(int(sm_ids[i]), round(float(per_sm[i]), 1)) for i in order[:12]])
        per_sm_all.append(dict(zip(sm_ids.tolist(), (per_sm / per_sm.mean()).tolist())))
        del Wb
    common = set(per_sm_all[0])
    for d in per_sm_all[1:]:
        common &= set(d)
    rel = np.array([[d[s] for s in sorted(common)] for d in per_sm_all])
    print("correlation of the per-SM relative stream time between shapes (1 = the same SMs are slow everywhere):")
    print(np.round(np.corrcoef(rel), 2))
    mean_rel = rel.mean(0)
    print("relative per-SM stream time averaged over shapes, by smid:")
    print(" ".join(f"{s}:{v:.2f}" for s, v in zip(sorted(common), mean_rel)))


if __name__ == "__main__":
    main()
