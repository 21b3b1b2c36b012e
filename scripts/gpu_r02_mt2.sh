#!/bin/bash
# Same-box A/B of the 512-row pair-tile rule: MT=1 everywhere / K >= 8192 (default) / K >= 2048.
mkdir -p gpurun_out
tag=${1:-r02t}
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs --new-tokens 16"
show() { python - "$1" <<PY
import json,sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "prefill_ms", round(d["phases_ms"]["prefill_ms"],1), "vit_ms", round(d["phases_ms"]["vit_ms"],2))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
for i in 1 2 3; do
  TEO_PAIR_MT=1 timeout 600 $B > gpurun_out/mt_${tag}_mt1_$i.json 2> /dev/null; show gpurun_out/mt_${tag}_mt1_$i.json
  timeout 600 $B > gpurun_out/mt_${tag}_k8192_$i.json 2> /dev/null; show gpurun_out/mt_${tag}_k8192_$i.json
  TEO_PAIR_MT=2048 timeout 600 $B > gpurun_out/mt_${tag}_k2048_$i.json 2> /dev/null; show gpurun_out/mt_${tag}_k2048_$i.json
done
