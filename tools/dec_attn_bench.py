"""Times the paged decode-attention kernel alone at the bench shape (bs=32, S≈2258, 32 heads × 128), rotating over
8 layer-sized KV pools so that no launch hits L2.  TEO_DEC_ATTN=cuda selects the persistent CUDA-core kernel, v1 the one-CTA-per-item kernel (A/B).
Development tool: `python tools/dec_attn_bench.py [B] [S]` on a B200."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from teochat_b200 import lib as L  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 2258
    H, hd, ps, iters, n_pools = 32, 128, 64, 40, 8
    import ctypes as C
    lib = L.load()
    h = C.c_void_p()
    L.check(lib.teo_create(0, C.byref(h)))
    dev = "cuda"
    pages_per = (S + ps - 1) // ps
    pool = torch.empty(n_pools, B * pages_per, 2, H, ps, hd, dtype=torch.bfloat16, device=dev)
    pool.view(-1)[: 1 << 20].normal_()
    bt = torch.arange(B * pages_per, dtype=torch.int32, device=dev).view(B, pages_per)
    q = torch.randn(B, 3 * H * hd, device=dev).to(torch.bfloat16)
    sl = torch.full((B,), S, dtype=torch.int32, device=dev)
    out = torch.empty(B, H * hd, dtype=torch.bfloat16, device=dev)
    ws = torch.empty(lib.teo_decode_attention_workspace_bytes(B, H, hd, 32), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def launch(i):
        L.check(lib.teo_decode_attention_h(h, q.data_ptr(), 3 * H * hd, pool[i % n_pools].data_ptr(), bt.data_ptr(), pages_per, sl.data_ptr(),
                                           out.data_ptr(), B, H, hd, ps, S, hd ** -0.5, ws.data_ptr(), ws.numel(), st))
    for i in range(4):
        launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        launch(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = B * 2 * H * hd * 2 * S
    print(f"decode attention [{os.environ.get('TEO_DEC_ATTN', 'mma')}] bs={B} S={S}: {ms * 1e3:.1f} us per launch (incl. combine), "
          f"{nbytes / ms / 1e6:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
