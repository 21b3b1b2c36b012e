#!/bin/bash
# Round-2 perf check 3: whole GPU suite on the current build, bench A/B over pair-GEMM L2 hints / group size, launch list.
mkdir -p gpurun_out
tag=${1:-r02e}
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_${tag}.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu_${tag}.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs"
show() { python - "$1" <<PY
import json,sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],1), "decode_ms", round(d["phases_ms"]["decode_ms"],1), "prefill_ms", round(d["phases_ms"]["prefill_ms"],1), "vit_ms", round(d["phases_ms"]["vit_ms"],2), "vit_fps", round(d["vit_frames_per_s"]))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
timeout 600 $B > gpurun_out/ab_${tag}_default.json 2> gpurun_out/ab_${tag}_default.err; show gpurun_out/ab_${tag}_default.json
TEO_PAIR_L2_HINT=2 timeout 600 $B > gpurun_out/ab_${tag}_hint2.json 2> gpurun_out/ab_${tag}_hint2.err; show gpurun_out/ab_${tag}_hint2.json
TEO_PAIR_GROUP_M=8 timeout 600 $B > gpurun_out/ab_${tag}_gm8.json 2> gpurun_out/ab_${tag}_gm8.err; show gpurun_out/ab_${tag}_gm8.json
TEO_PAIR_GROUP_M=8 TEO_PAIR_L2_HINT=2 timeout 600 $B > gpurun_out/ab_${tag}_gm8hint2.json 2> gpurun_out/ab_${tag}_gm8hint2.err; show gpurun_out/ab_${tag}_gm8hint2.json
timeout 600 $B > gpurun_out/ab_${tag}_default2.json 2> gpurun_out/ab_${tag}_default2.err; show gpurun_out/ab_${tag}_default2.json
if [ "$2" == "list" ]; then
  timeout 900 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --profile --warmup 1 --new-tokens 5 > gpurun_out/prof_launch_${tag}.log 2>&1; echo "launchlist exit $?"
fi
