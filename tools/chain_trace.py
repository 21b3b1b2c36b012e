"""Timeline of the persistent decode chain kernel (decode_chain.cu) from its in-kernel %globaltimer stamps (development hook
teo_dbg_chain_trace): where a layer's four GEMM phases spend their time — weight streaming, waiting for the grid barriers, the
fused reductions.  `python tools/chain_trace.py [batch] [frames] [l2_prefetch]` on a B200 (full-size model, eager decode steps)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from teochat_b200 import lib as L  # noqa: E402
from teochat_b200.config import TeoConfig  # noqa: E402
from teochat_b200.engine import TeoModel  # noqa: E402
from teochat_b200.weights import TeoWeights  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    pf = int(sys.argv[3]) if len(sys.argv) > 3 else None
    dev = torch.device("cuda", 0)
    cfg = TeoConfig.full()
    model = TeoModel(cfg, TeoWeights.from_synthetic(cfg, 1234, dev), dev)
    raw = C.CDLL(L.lib_path())
    raw.teo_dbg_chain_trace.restype = C.c_longlong
    raw.teo_dbg_chain_trace.argtypes = [C.c_void_p, C.c_int]
    if pf is not None:
        raw.teo_dbg_chain_prefetch(C.c_int(pf))
    wl = bench.Workload(model, cfg, 0, 1, dev, T, B, 6)
    model.use_graph = False
    wl.step(wl.dev_list)                                    # warm
    n_launch = 64
    buf = torch.zeros(n_launch, 148, 4, 8, dtype=torch.int64, device=dev)
    raw.teo_dbg_chain_trace(buf.data_ptr(), n_launch)
    wl.step(wl.dev_list)
    torch.cuda.synchronize()
    n = raw.teo_dbg_chain_trace(None, 0)
    t = buf.cpu().numpy().astype(np.float64)
    print(f"batch {B}, T={T}, l2_prefetch {pf}: {n} chain launches stamped (ring of {n_launch})")
    use = [i for i in range(min(n, n_launch)) if t[i, :, 3, 3].min() > 0]          # 4-phase launches only
    names = ["o_proj", "gate_up", "down", "qkv/lm_head"]
    for p in range(4):
        rows = []
        for i in use:
            x = t[i, :, p, :]
            start = t[i, :, 0, 0].min()
            rows.append([
                (x[:, 1] - x[:, 0]).mean(),                       # L2 prefetch issue
                (x[:, 2] - x[:, 1]).mean(),                       # producer waits for the input
                x[:, 2].max() - start,                            # input ready (since kernel start)
                (x[:, 3] - x[:, 2]).mean(),                       # stream + MMA + partial stores
                (x[:, 3].max() - x[:, 3].min()),                  # skew of the CTAs' finish times
                (x[:, 4] - x[:, 3]).mean(),                       # barrier "partials complete"
                (x[:, 5] - x[:, 4]).mean(),                       # reduction
                x[:, 5].max() - start,                            # phase done (since kernel start)
            ])
        r = np.median(np.array(rows), axis=0) / 1e3
        print(f"  phase {p} {names[p]:12s} pf-issue {r[0]:5.1f}  wait-input {r[1]:5.1f}  [input ready @{r[2]:6.1f}]  stream {r[3]:5.1f}  finish-skew {r[4]:5.1f}  "
              f"barrier {r[5]:5.1f}  reduce {r[6]:5.1f}  [done @{r[7]:6.1f}] us")
    tot = np.median([t[i, :, 3, 5].max() - t[i, :, 0, 0].min() for i in use]) / 1e3
    print(f"  chain launch, first CTA start → last reduce done: {tot:.1f} us (median of {len(use)} launches)")


if __name__ == "__main__":
    main()
