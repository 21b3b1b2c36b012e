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
    assert bench.pick_cpu_threads() == n                # fixed policy: every usable CPU


def test_configs_table_matches_baseline_json():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        cfgs = json.load(f)["configs"]
    assert "bs=64" in cfgs[1] and "128 output tokens" in cfgs[1] and bench.CONFIGS[1] == (1, 64, 128)
    assert "T=8" in cfgs[2] and "bs=32" in cfgs[2] and "256 output tokens" in cfgs[2] and bench.CONFIGS[2] == (8, 32, 256)
    assert "T=16" in cfgs[4] and "bs=16" in cfgs[4] and "512 output tokens" in cfgs[4] and bench.CONFIGS[4] == (16, 16 // 8, 512)
    assert bench.CPU_SAMPLE == (2, 16) and "2-frame" in cfgs[0] and "16 tokens" in cfgs[0]


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
    assert line["same_config"] is False and line["kind"] == "port" and "configs[0]" in cb["sample"] and "configs[0]" in line["config"]["sample"]
    assert "configs[2]" in line["config"]["workload"]            # the GPU arm's workload is named, the CPU sample beside it
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
