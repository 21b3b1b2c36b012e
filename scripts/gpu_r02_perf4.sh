#!/bin/bash
# Round-2 perf check 4: deeper stream-K ring — GEMM-level tests, bench, decode GEMM timeline.
mkdir -p gpurun_out
tag=${1:-r02h}
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu_${tag}.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu_${tag}.log
timeout 400 python tools/dec_gemm_skew.py 32 > gpurun_out/dec_gemm_skew_${tag}.log 2>&1; echo "skew exit $?"; grep -E "kernel span" gpurun_out/dec_gemm_skew_${tag}.log | cut -c1-200
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
timeout 600 $B > gpurun_out/ab_${tag}_default.json 2> gpurun_out/ab_${tag}_default.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/ab_${tag}_default.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["phases_ms"], "vit_fps", round(d["vit_frames_per_s"]))
for k, v in d.get("other_configs", {}).items():
    print(k, round(v["value"],1), v["phases_ms"])
PY
