#!/bin/bash
# Folded LayerNorm in the ViT: op test, the model-level parity tests that cover the tower, bench A/B.
mkdir -p gpurun_out
tag=${1:-r02j}
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_exact.py -m gpu -q -s -p no:cacheprovider --timeout 600 -k "folded or vit or reference_code or full_width or config1 or benchmark_contexts or synthetic_weights or drop_in or hf_directory" > gpurun_out/pytest_ln_${tag}.log 2>&1; echo "tests exit $?"; grep -E "passed|failed|rel err|FAILED|Error" gpurun_out/pytest_ln_${tag}.log | cut -c1-260 | head -40
for f in 1 0 1 0; do
  TEO_VIT_LN_FOLD=$f timeout 300 python bench.py --steps 3 --warmup 2 --new-tokens 8 --no-cpu-baseline --no-other-configs > gpurun_out/ln_${tag}_f$f.json 2> gpurun_out/ln_${tag}_f$f.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ln_${tag}_f$f.json").read().strip().splitlines()[-1])
    print("fold $f: vit_ms", round(d["phases_ms"]["vit_ms"], 2), "vit_fps", round(d["vit_frames_per_s"]), "prefill_ms", round(d["phases_ms"]["prefill_ms"], 1))
except Exception as e:
    print("fold $f: no line", e)
PY
done
