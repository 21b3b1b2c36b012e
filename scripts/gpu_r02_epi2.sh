#!/bin/bash
# Same-box A/B of the epilogue staging depth: variant lib built with -DTEO_EPI_BUFS=1 (one 4 KB staging buffer per warp, rings 6 / 4 stages)
# against the tree's default (two buffers, rings 5 / 3), each under the K >= 8192 rule and the K >= 2048 rule of the 512-row tiles.
mkdir -p gpurun_out
tag=${1:-r02u}
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or pair" 2>&1 | tail -3
TEO_PAIR_MT=2 timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or pair" 2>&1 | tail -3
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
V=teochat_b200/lib/variants/epi1.so
for i in 1 2 3; do
  TEO_LIB_PATH=$V timeout 600 $B > gpurun_out/epi_${tag}_b1_k8192_$i.json 2> /dev/null; show gpurun_out/epi_${tag}_b1_k8192_$i.json
  timeout 600 $B > gpurun_out/epi_${tag}_b2_k8192_$i.json 2> /dev/null; show gpurun_out/epi_${tag}_b2_k8192_$i.json
  TEO_PAIR_MT=2048 timeout 600 $B > gpurun_out/epi_${tag}_b2_k2048_$i.json 2> /dev/null; show gpurun_out/epi_${tag}_b2_k2048_$i.json
  TEO_PAIR_MT=1 timeout 600 $B > gpurun_out/epi_${tag}_b2_mt1_$i.json 2> /dev/null; show gpurun_out/epi_${tag}_b2_mt1_$i.json
done
