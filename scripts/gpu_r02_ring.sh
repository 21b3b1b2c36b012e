#!/bin/bash
# A/B of the stream-K ring depth (default 11/9/7 stages vs the round-1 8/8/6 + staging: lib/variants/sk8.so), same box, alternating; then the default bench line.
mkdir -p gpurun_out
tag=${1:-r02m}
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs"
show() { python - "$1" <<PY
import json,sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], "value", round(d["value"],1), "decode_ms", round(d["phases_ms"]["decode_ms"],1), "prefill_ms", round(d["phases_ms"]["prefill_ms"],1), "roofline", round(r.get("frac", 0), 3), r.get("us_per_launch_blocks"))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
for i in 1 2 3; do
  timeout 600 $B > gpurun_out/ring_${tag}_deep$i.json 2> gpurun_out/ring_${tag}_deep$i.err; show gpurun_out/ring_${tag}_deep$i.json
  TEO_LIB_PATH=teochat_b200/lib/variants/sk8.so timeout 600 $B > gpurun_out/ring_${tag}_sk8_$i.json 2> gpurun_out/ring_${tag}_sk8_$i.err; show gpurun_out/ring_${tag}_sk8_$i.json
done
for c in 1 4; do
  timeout 600 $B --config $c > gpurun_out/ring_${tag}_c${c}_deep.json 2> /dev/null; show gpurun_out/ring_${tag}_c${c}_deep.json
  TEO_LIB_PATH=teochat_b200/lib/variants/sk8.so timeout 600 $B --config $c > gpurun_out/ring_${tag}_c${c}_sk8.json 2> /dev/null; show gpurun_out/ring_${tag}_c${c}_sk8.json
done
