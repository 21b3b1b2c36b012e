#!/bin/bash
# ViT chunking A/B (frames per teo_vit_encode call: L2 residency of the LayerNorm output vs wave quantisation of the GEMMs).
mkdir -p gpurun_out
tag=${1:-r02i}
for c in 512 128 64 32 512; do
  TEO_VIT_CHUNK=$c timeout 300 python bench.py --steps 3 --warmup 2 --new-tokens 8 --no-cpu-baseline --no-other-configs > gpurun_out/vit_${tag}_c$c.json 2> gpurun_out/vit_${tag}_c$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/vit_${tag}_c$c.json").read().strip().splitlines()[-1])
    print("chunk $c: vit_ms", round(d["phases_ms"]["vit_ms"], 2), "vit_fps", round(d["vit_frames_per_s"]), "prefill_ms", round(d["phases_ms"]["prefill_ms"], 1))
except Exception as e:
    print("chunk $c: no line", e)
PY
done
