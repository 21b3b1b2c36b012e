#!/bin/bash
# Same-box A/B of the whole step: the build this session started from (commit 59d6354 compiled as lib/variants/r02start.so) against the tree's build,
# alternating, default bench workload with 64 new tokens (ViT and prefill are unaffected by the token count).
mkdir -p gpurun_out
tag=${1:-r02w}
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs --new-tokens 64"
for i in 1 2 3; do
  for v in r02start ""; do
    if [ -z "$v" ]; then name=new; unset TEO_LIB_PATH; else name=$v; export TEO_LIB_PATH=teochat_b200/lib/variants/$v.so; fi
    timeout 600 $B > gpurun_out/ab_${tag}_${name}_$i.json 2> /dev/null
    python - gpurun_out/ab_${tag}_${name}_$i.json <<PY | tee -a gpurun_out/ab_${tag}.log
import json,sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "vit_ms", round(d["phases_ms"]["vit_ms"],2), "prefill_ms", round(d["phases_ms"]["prefill_ms"],1), "decode_ms/step", round(d["phases_ms"]["decode_ms"]/63,3), "vit_frames_per_s", round(d["vit_frames_per_s"]))
PY
  done
done
