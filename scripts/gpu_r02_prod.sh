#!/bin/bash
# A/B of one vs two TMA producer warps in the stream-K decode GEMM (TEO_SK_PRODUCERS), same box, alternating runs; correctness first.
mkdir -p gpurun_out
tag=${1:-r02p}
TEO_SK_PRODUCERS=2 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider --timeout 600 -k "gemm or generate or graph or decode_chain or golden or full_width" > gpurun_out/pytest_prod_${tag}.log 2>&1; echo "tests (2 producers) exit $?"; tail -n 3 gpurun_out/pytest_prod_${tag}.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs"
show() { python - "$1" <<PY
import json,sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],1), "decode_ms", round(d["phases_ms"]["decode_ms"],1), "prefill_ms", round(d["phases_ms"]["prefill_ms"],1))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
for i in 1 2 3; do
  TEO_SK_PRODUCERS=1 timeout 600 $B > gpurun_out/prod_${tag}_p1_$i.json 2> gpurun_out/prod_${tag}_p1_$i.err; show gpurun_out/prod_${tag}_p1_$i.json
  TEO_SK_PRODUCERS=2 timeout 600 $B > gpurun_out/prod_${tag}_p2_$i.json 2> gpurun_out/prod_${tag}_p2_$i.err; show gpurun_out/prod_${tag}_p2_$i.json
done
for c in 1 4; do
  TEO_SK_PRODUCERS=1 timeout 600 $B --config $c > gpurun_out/prod_${tag}_c${c}_p1.json 2> /dev/null; show gpurun_out/prod_${tag}_c${c}_p1.json
  TEO_SK_PRODUCERS=2 timeout 600 $B --config $c > gpurun_out/prod_${tag}_c${c}_p2.json 2> /dev/null; show gpurun_out/prod_${tag}_c${c}_p2.json
done
TEO_SK_PRODUCERS=2 timeout 300 python tools/dec_gemm_skew.py 32 > gpurun_out/dec_gemm_skew_${tag}_p2.log 2>&1; grep "kernel span" gpurun_out/dec_gemm_skew_${tag}_p2.log | cut -c1-180
