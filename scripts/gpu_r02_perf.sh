#!/bin/bash
# Round-2 perf check on one B200: decode-chain bit-identity tests, serving tests, chain on/off bench A/B, pair-GEMM sweep (+ ncu dram bytes).
mkdir -p gpurun_out
tag=${1:-r02c}
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s -p no:cacheprovider --timeout 600 -k "decode_chain" > gpurun_out/pytest_chain_${tag}.log 2>&1; echo "chain tests exit $?"; grep -E "passed|failed|kernels per decode|Error|error|diff" gpurun_out/pytest_chain_${tag}.log | head -20
timeout 900 python -m pytest tests/test_gpu_serving.py tests/test_gpu_exact.py -m gpu -q -p no:cacheprovider --timeout 600 -k "not benchmark_contexts and not config1" > gpurun_out/pytest_serving_${tag}.log 2>&1; echo "serving tests exit $?"; tail -n 15 gpurun_out/pytest_serving_${tag}.log
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_chain_${tag}.json 2> gpurun_out/bench_chain_${tag}.err; echo "bench(chain) exit $?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_chain_${tag}.json").read().strip().splitlines()[-1])
    print("chain:", {k: d[k] for k in ("value", "ms_per_step", "phases_ms", "decode_tokens_per_s")}, {k: (v["value"], v["phases_ms"]) for k, v in d.get("other_configs", {}).items()})
except Exception as e:
    print("no chain bench line", e)
PY
TEO_DEC_CHAIN=0 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs > gpurun_out/bench_nochain_${tag}.json 2> gpurun_out/bench_nochain_${tag}.err; echo "bench(no chain) exit $?"; cut -c1-900 gpurun_out/bench_nochain_${tag}.json
if [ "$2" == "sweep" ]; then
  timeout 600 python tools/pair_sweep.py time > gpurun_out/pair_sweep_${tag}.log 2>&1; echo "sweep exit $?"; cat gpurun_out/pair_sweep_${tag}.log
  for cfg in "16 1048576 0 0" "16 1048576 0 1" "24 24 0 1" "16 16 1 1"; do
    timeout 300 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:gemm_pair_kernel --csv python tools/pair_sweep.py one $cfg > "gpurun_out/pair_ncu_${tag}_$(echo $cfg | tr ' ' '_').csv" 2>&1; echo "ncu $cfg exit $?"
  done
fi
