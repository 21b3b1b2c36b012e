"""In-situ timeline of the decode GEMMs: runs the real decode loop (eager launches with PDL, TEO_NO_GRAPH=1) at the bench shape
and prints, for the last decode step, each GEMM's phases relative to the previous GEMM's exit (development hook
teo_dbg_gemm_trace; %globaltimer stamps).  `python tools/dec_step_trace.py [frames] [batch]` on a B200."""
import ctypes as C
import os
import sys

os.environ["TEO_NO_GRAPH"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from teochat_b200.config import TeoConfig  # noqa: E402
from teochat_b200.engine import TeoModel  # noqa: E402
from teochat_b200.weights import TeoWeights  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    cfg = TeoConfig.full()
    dev = torch.device("cuda:0")
    model = TeoModel(cfg, TeoWeights.from_synthetic(cfg, 1234, dev), dev)
    lib = model.lib
    lib.teo_dbg_gemm_trace.restype = C.c_longlong
    lib.teo_dbg_gemm_trace.argtypes = [C.c_void_p, C.c_int]
    ids = [bench.make_prompt_ids(cfg, T) for _ in range(B)]
    frames = [torch.randint(0, 256, (T, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(B)]
    model.generate_batch(ids, frames_u8=frames, max_new_tokens=6, eos_token_id=-1)           # warm-up
    per_step = cfg.llama.num_hidden_layers * 4 + 1
    cap = per_step * 2
    buf = torch.zeros(cap * 148 * 8, dtype=torch.int64, device=dev)
    lib.teo_dbg_gemm_trace(buf.data_ptr(), cap)
    model.generate_batch(ids, frames_u8=frames, max_new_tokens=8, eos_token_id=-1)
    torch.cuda.synchronize()
    n = lib.teo_dbg_gemm_trace(None, 0)
    t = buf.view(cap, 148, 8).cpu()
    order = [(n - per_step + j) % cap for j in range(per_step)]                               # the last decode step, in launch order
    names = ["qkv", "o", "gate_up", "down"]
    prev_exit = None
    rows = []
    for j, slot in enumerate(order):
        x = t[slot]
        x = x[x[:, 0] > 0].double()
        ent, ring, dep, first, last, epi, ext = (x[:, c] for c in (0, 2, 3, 4, 5, 6, 7))
        e0 = ent.min()
        rows.append((names[j % 4] if j < per_step - 1 else "lm_head", j // 4,
                     None if prev_exit is None else (e0 - prev_exit) / 1e3,           # first CTA enters, relative to previous GEMM's last exit
                     (ent.max() - e0) / 1e3, (ring.median() - e0) / 1e3, (dep.median() - e0) / 1e3, (first.median() - e0) / 1e3,
                     (last.median() - e0) / 1e3, (ext.max() - e0) / 1e3))
        prev_exit = ext.max()
    print("GEMM      layer  enter-prev_exit  last_enter  ring_filled  dep_done  first_ops  last_mma  exit   (us; columns 4+ from this GEMM's first entry)")
    for r in rows[8:24]:
        print(f"{r[0]:8s} {r[1]:5d}  {'' if r[2] is None else '%8.2f' % r[2]:>15s}  {r[3]:10.2f}  {r[4]:11.2f}  {r[5]:8.2f}  {r[6]:9.2f}  {r[7]:8.2f}  {r[8]:6.2f}")
    import statistics
    for nm in names:
        sel = [r for r in rows[4:100] if r[0] == nm]
        print(f"{nm:8s} median: gap after previous GEMM exit {statistics.median(r[2] for r in sel):6.2f} us, dep_done {statistics.median(r[5] for r in sel):6.2f}, "
              f"first_ops {statistics.median(r[6] for r in sel):6.2f}, span {statistics.median(r[8] for r in sel):6.2f}")
    step_us = (t[order[-1]][:, 7].max() - t[order[0]][:, 0][t[order[0]][:, 0] > 0].min()).item() / 1e3
    print(f"first qkv entry → lm_head exit: {step_us:.1f} us")


if __name__ == "__main__":
    main()
