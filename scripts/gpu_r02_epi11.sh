#!/bin/bash
# Compile-time-specialised chunk conversion in the staged epilogue: GEMM tests (both tile rules), per-tile timeline, sustained rates against
# cuBLAS, isolated-launch tensor-pipe activity, bench phases for the 512-row rule K >= 4096 / K >= 1024.
mkdir -p gpurun_out
tag=${1:-r02u}
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or pair or swiglu or linear or layernorm" -p no:cacheprovider 2>&1 | tail -2
TEO_PAIR_MT=1024 timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or pair or swiglu or linear or layernorm" -p no:cacheprovider 2>&1 | tail -2
TEO_LIB_PATH=teochat_b200/lib/variants/pairtrace.so timeout 300 python tools/pair_trace.py prefill 2>&1 | grep -E "===|tile 4:|tile period" | cut -c1-260 | tee gpurun_out/pair_trace_prefill_${tag}.log
TEO_LIB_PATH=teochat_b200/lib/variants/pairtrace.so TEO_PAIR_MT=1024 timeout 300 python tools/pair_trace.py vit 2>&1 | grep -E "===|tile 4:|tile period" | cut -c1-260 | tee gpurun_out/pair_trace_vit_mt2_${tag}.log
TEO_LIB_PATH=teochat_b200/lib/variants/pairtrace.so timeout 300 python tools/pair_trace.py vit 2>&1 | grep -E "===|tile 4:|tile period" | cut -c1-260 | tee gpurun_out/pair_trace_vit_mt1_${tag}.log
for mt in "" 1024; do
  for w in prefill vit; do
    TEO_PAIR_MT=$mt timeout 600 python tools/pair_sweep.py cublas $w 2>&1 | grep "ours/cuBLAS" | sed "s/^/MT=[$mt] /" | tee -a gpurun_out/epi11_${tag}.log
  done
done
bash scripts/gpu_r02_epi8.sh ${tag} 2>&1 | tee -a gpurun_out/epi11_${tag}.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs --new-tokens 16"
for i in 1 2; do
  for mt in 4096 1024; do
    TEO_PAIR_MT=$mt timeout 600 $B > gpurun_out/epi11_${tag}_mt${mt}_$i.json 2> /dev/null
    python - gpurun_out/epi11_${tag}_mt${mt}_$i.json <<PY | tee -a gpurun_out/epi11_${tag}.log
import json,sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "prefill_ms", round(d["phases_ms"]["prefill_ms"],1), "vit_ms", round(d["phases_ms"]["vit_ms"],2))
PY
  done
done
