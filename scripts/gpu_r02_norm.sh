#!/bin/bash
# Norm kernels specialised on the vectors a thread holds (4 instead of 8): op tests, model tests, launch-list times, bench phases.
mkdir -p gpurun_out
tag=${1:-r02nm}
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q -m gpu -k "norm or vit or golden or tiny or full_width" -p no:cacheprovider 2>&1 | tail -2
bash scripts/gpu_launchlist.sh 2 ${tag} | grep -E "norm_kernel|ViT \+ projector|prefill \(" 
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs --new-tokens 16"
for i in 1 2; do timeout 600 $B 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('vit_ms', round(d['phases_ms']['vit_ms'],2), 'prefill_ms', round(d['phases_ms']['prefill_ms'],1), 'vit_fps', round(d['vit_frames_per_s']))"; done
