#!/bin/bash
# Pipelined-residual / fast-SwiGLU epilogue: GEMM tests, then gemm_pair_kernel against cuBLAS on the same operands for the 512-row
# tile rule K >= 8192 (default) / K >= 4096 / K >= 1024 / never (TEO_PAIR_MT), then the bench phases for the same rules.
mkdir -p gpurun_out
tag=${1:-r02f}
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or pair or swiglu or linear" -p no:cacheprovider 2>&1 | tail -3
TEO_PAIR_MT=1024 timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "gemm or pair or swiglu or linear" -p no:cacheprovider 2>&1 | tail -3
for mt in "" 4096 1024 1; do
  for w in prefill vit; do
    TEO_PAIR_MT=$mt timeout 600 python tools/pair_sweep.py cublas $w 2>&1 | grep "ours/cuBLAS" | sed "s/^/MT=[$mt] /" | tee -a gpurun_out/epi3_${tag}.log
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
for i in 1 2; do
  for mt in 8192 4096 1024; do
    TEO_PAIR_MT=$mt timeout 600 $B > gpurun_out/epi3_${tag}_mt${mt}_$i.json 2> /dev/null; show gpurun_out/epi3_${tag}_mt${mt}_$i.json | tee -a gpurun_out/epi3_${tag}.log
  done
done
# decode attention streaming skeleton: default / compute skipped / compute skipped + linear 16 KiB bulk copies
for v in "" 1 2; do
  TEO_DEC_DBG_SKIP=$v timeout 300 python tools/dec_attn_bench.py 2>&1 | tail -n 1 | sed "s/^/skip=[$v] /" | tee -a gpurun_out/dec_attn_${tag}.log
done
