#!/bin/bash
# Round-2 perf check 2: chain v2 (bit-identity, timeline, A/B against the per-GEMM path with 3 L2-prefetch depths), pair-GEMM sweep (round-robin).
mkdir -p gpurun_out
tag=${1:-r02d}
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -s -p no:cacheprovider --timeout 600 -k "decode_chain or graph_replay" > gpurun_out/pytest_chain_${tag}.log 2>&1; echo "chain tests exit $?"; grep -E "passed|failed|kernels per decode|rror|diff" gpurun_out/pytest_chain_${tag}.log | head -20
for pf in 24 0; do timeout 300 python tools/chain_trace.py 32 8 $pf > gpurun_out/chain_trace_${tag}_pf${pf}.log 2>&1; echo "trace pf=$pf exit $?"; cat gpurun_out/chain_trace_${tag}_pf${pf}.log; done
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-other-configs"
show() { python - "$1" <<PY
import json,sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],1), "decode_ms", round(d["phases_ms"]["decode_ms"],1), "prefill_ms", round(d["phases_ms"]["prefill_ms"],1), "vit_ms", round(d["phases_ms"]["vit_ms"],2))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
TEO_DEC_CHAIN=0 timeout 600 $B > gpurun_out/ab_${tag}_nochain.json 2> gpurun_out/ab_${tag}_nochain.err; show gpurun_out/ab_${tag}_nochain.json
for pf in 0 24 48 -24; do TEO_CHAIN_PF=$pf timeout 600 $B > gpurun_out/ab_${tag}_pf${pf}.json 2> gpurun_out/ab_${tag}_pf${pf}.err; show gpurun_out/ab_${tag}_pf${pf}.json; done
TEO_DEC_CHAIN=0 timeout 600 $B --config 4 > gpurun_out/ab_${tag}_c4_nochain.json 2> /dev/null; show gpurun_out/ab_${tag}_c4_nochain.json
TEO_CHAIN_PF=24 timeout 600 $B --config 4 > gpurun_out/ab_${tag}_c4_pf24.json 2> /dev/null; show gpurun_out/ab_${tag}_c4_pf24.json
if [ "$2" == "sweep" ]; then
  timeout 600 python tools/pair_sweep.py time prefill > gpurun_out/pair_sweep_${tag}.log 2>&1; echo "sweep exit $?"; cat gpurun_out/pair_sweep_${tag}.log
  timeout 300 python tools/pair_sweep.py time vit > gpurun_out/pair_sweep_vit_${tag}.log 2>&1; echo "sweep vit exit $?"; cat gpurun_out/pair_sweep_vit_${tag}.log
  for cfg in "16 1048576 0 2" "32 1048576 0 2" "8 1048576 0 0" "32 1048576 0 0"; do
    timeout 300 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:gemm_pair_kernel --csv python tools/pair_sweep.py one $cfg > "gpurun_out/pair_ncu_${tag}_$(echo $cfg | tr ' ' '_').csv" 2>&1; echo "ncu $cfg exit $?"
  done
fi
