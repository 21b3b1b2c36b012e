"""World-size-2 gloo test of the data-parallel plumbing: contiguous sharding and
the single end-of-run all-gather reproduce dataset order."""
import os
import socket
import subprocess
import sys
import textwrap

from teochat_b200.dist import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for n in (0, 1, 7, 32, 256):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_tokens_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys, json
        sys.path.insert(0, {ROOT!r})
        import torch
        from teochat_b200 import dist as TD
        rank, world, _ = TD.init_from_env("gloo")
        n_total, max_new = 5, 6
        lo, hi = TD.shard_range(n_total, rank, world)
        outs = [[100 * i + j for j in range(1 + i % max_new)] for i in range(lo, hi)]     # "generated ids" of example i
        rows = max(h - l for l, h in (TD.shard_range(n_total, r, world) for r in range(world)))
        got = TD.gather_tokens(TD.pack_tokens(outs, rows, max_new, "cpu"), n_total)
        want = [[100 * i + j for j in range(1 + i % max_new)] for i in range(n_total)]
        assert got == want, (got, want)
        if rank == 0:
            print("OK")
    """))
    for attempt in range(3):           # a just-released port can be taken by another process: retry on a fresh one
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=240)
        if r.returncode == 0 and "OK" in r.stdout:
            return
        if "AssertionError" in r.stderr:   # the check itself failed: no point retrying
            break
    assert False, r.stdout + r.stderr
