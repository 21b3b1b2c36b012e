#!/bin/bash
# decode-step A/B inside the real (graph + PDL) decode loop: each argument is "name:ENV=VAL,ENV=VAL" (or "name:")
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  ( IFS=,; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    timeout 900 python bench.py --no-cpu-baseline --steps 1 --warmup 1 --new-tokens 128 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_$name.json'))
    print('$name', 'decode tok/s %.0f' % d['decode_tokens_per_s'], 'decode ms/step %.3f' % (d['phases_ms']['decode_ms']/127), 'prefill ms %.1f' % d['phases_ms']['prefill_ms'], 'vit ms %.2f' % d['phases_ms']['vit_ms'], 'value %.0f' % d['value'])
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/ab_$name.err').read()[-600:])
PY
  )
done
