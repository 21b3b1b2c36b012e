#!/bin/bash
# Epilogue bias hoist: GEMM tests + ViT tests, ViT timing (bench with short decode), cuBLAS comparison at the ViT shapes.
mkdir -p gpurun_out
tag=${1:-r02k}
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider --timeout 600 > gpurun_out/pytest_${tag}.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/pytest_${tag}.log; grep -E "folded LayerNorm \{|ViT\+projector vs" gpurun_out/pytest_${tag}.log | head
for i in 1 2; do
  timeout 300 python bench.py --steps 3 --warmup 2 --new-tokens 8 --no-cpu-baseline --no-other-configs > gpurun_out/epi_${tag}_$i.json 2> gpurun_out/epi_${tag}_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/epi_${tag}_$i.json").read().strip().splitlines()[-1])
    print("run $i: vit_ms", round(d["phases_ms"]["vit_ms"], 2), "vit_fps", round(d["vit_frames_per_s"]), "prefill_ms", round(d["phases_ms"]["prefill_ms"], 1))
except Exception as e:
    print("run $i: no line", e)
PY
done
timeout 300 python tools/pair_sweep.py cublas vit > gpurun_out/pair_vs_cublas_${tag}.log 2>&1; echo "cublas vit exit $?"; cat gpurun_out/pair_vs_cublas_${tag}.log
