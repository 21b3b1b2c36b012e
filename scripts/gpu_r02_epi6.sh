#!/bin/bash
# Epilogue warps gated by a named barrier (one polling warp, suspend-time hint) against the all-warps-poll variant, same box:
# GEMM tests, sustained GEMM rates against cuBLAS, bench phases.
mkdir -p gpurun_out
tag=${1:-r02i}
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or pair or swiglu or linear" -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do
  for v in "" teochat_b200/lib/variants/pollall.so; do
    for w in prefill vit; do
      TEO_LIB_PATH=$v timeout 600 python tools/pair_sweep.py cublas $w 2>&1 | grep "ours/cuBLAS" | sed "s|^|lib=[$v] |" | tee -a gpurun_out/epi6_${tag}.log
    done
  done
done
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
  timeout 600 $B > gpurun_out/epi6_${tag}_gate_$i.json 2> /dev/null; show gpurun_out/epi6_${tag}_gate_$i.json | tee -a gpurun_out/epi6_${tag}.log
  TEO_LIB_PATH=teochat_b200/lib/variants/pollall.so timeout 600 $B > gpurun_out/epi6_${tag}_poll_$i.json 2> /dev/null; show gpurun_out/epi6_${tag}_poll_$i.json | tee -a gpurun_out/epi6_${tag}.log
done
