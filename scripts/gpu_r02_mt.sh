#!/bin/bash
# 512-row pair tiles (gemm_pair_kernel<2>): correctness (GEMM tests also with the kernel forced on small shapes), then same-box A/B of the bench.
mkdir -p gpurun_out
tag=${1:-r02s}
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider --timeout 600 -k "gemm" > gpurun_out/pytest_mt_${tag}.log 2>&1; echo "gemm tests exit $?"; tail -n 3 gpurun_out/pytest_mt_${tag}.log
TEO_PAIR_MT=2 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider --timeout 600 -k "gemm or vit or full_width or reference_code or golden" > gpurun_out/pytest_mt2_${tag}.log 2>&1; echo "tests with MT=2 forced exit $?"; tail -n 3 gpurun_out/pytest_mt2_${tag}.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs --new-tokens 16"
show() { python - "$1" <<PY
import json,sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "prefill_ms", round(d["phases_ms"]["prefill_ms"],1), "vit_ms", round(d["phases_ms"]["vit_ms"],2), "decode_ms", round(d["phases_ms"]["decode_ms"],1))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
for i in 1 2 3; do
  TEO_PAIR_MT=1 timeout 600 $B > gpurun_out/mt_${tag}_mt1_$i.json 2> gpurun_out/mt_${tag}_mt1_$i.err; show gpurun_out/mt_${tag}_mt1_$i.json
  timeout 600 $B > gpurun_out/mt_${tag}_auto_$i.json 2> gpurun_out/mt_${tag}_auto_$i.err; show gpurun_out/mt_${tag}_auto_$i.json
done
timeout 300 python tools/pair_sweep.py cublas prefill > gpurun_out/pair_vs_cublas_${tag}.log 2>&1; cat gpurun_out/pair_vs_cublas_${tag}.log
