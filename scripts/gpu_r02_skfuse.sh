#!/bin/bash
# In-kernel reduction + SwiGLU of the decode gate/up GEMM (TEO_SK_FUSE=1): bit-identity against the glue-kernel path, then a same-box A/B of the decode step.
mkdir -p gpurun_out
tag=${1:-r02y}
TEO_SK_FUSE=0 timeout 600 python tools/sk_fuse_check.py dump /tmp/skf0.pt 2>&1 | tail -5
TEO_SK_FUSE=1 timeout 600 python tools/sk_fuse_check.py dump /tmp/skf1.pt 2>&1 | tail -5
timeout 100 python tools/sk_fuse_check.py cmp /tmp/skf0.pt /tmp/skf1.pt 2>&1 | tail -6 | tee gpurun_out/skfuse_${tag}.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs --new-tokens 96"
for i in 1 2 3; do
  for f in 0 1; do
    TEO_SK_FUSE=$f timeout 600 $B > gpurun_out/skfuse_${tag}_f${f}_$i.json 2> /dev/null
    python - gpurun_out/skfuse_${tag}_f${f}_$i.json <<PY | tee -a gpurun_out/skfuse_${tag}.log
import json,sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "decode_ms/step", round(d["phases_ms"]["decode_ms"]/95,4), "launches", d["gpu_launches"])
PY
  done
done
