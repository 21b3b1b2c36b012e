"""bench.py host-side pieces that run without a GPU: the prompt of the benchmark workload, the CPU legs' thread choice and the
JSON line of the reference arm (checked on the tiny config so it stays a seconds-long test)."""
import json
import os
import subprocess
import sys

import bench
from teochat_b200.config import TeoConfig
from teochat_b200.constants import IMAGE_TOKEN_INDEX

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_benchmark_prompt_shape():
    cfg = TeoConfig.full()
    ids = bench.make_prompt_ids(cfg, 8)
    assert ids[0] == 1 and ids.count(IMAGE_TOKEN_INDEX) == 8
    s0 = len(ids) - 8 + 8 * cfg.tokens_per_image
    assert s0 == 2130                                  # the context length quoted in DESIGN.md / the bench line
    assert all(t == IMAGE_TOKEN_INDEX or 0 <= t < cfg.llama.vocab_size for t in ids)


def test_cpu_thread_choice_is_bounded():
    n = bench.host_cpus()
    assert 1 <= n <= (os.cpu_count() or 1)
    os.environ["TEO_CPU_THREADS"] = "3"
    try:
        assert bench.pick_cpu_threads() == 3
    finally:
        del os.environ["TEO_CPU_THREADS"]
    assert 1 <= bench.pick_cpu_threads() <= max(n, 4)


def test_reference_arm_json_line_on_tiny_config(monkeypatch, capsys):
    monkeypatch.setattr(TeoConfig, "full", staticmethod(TeoConfig.tiny))
    monkeypatch.setenv("TEO_CPU_THREADS", "2")
    args = bench.argparse.Namespace(gpus=1, steps=1, warmup=0, frames=8, batch=32, new_tokens=256)
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] == line["value"] and "new tokens" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
