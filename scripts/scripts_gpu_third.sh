#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 python -m pytest -q -m gpu -p no:cacheprovider --timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; }
run ops tests/test_gpu_ops.py
run model tests/test_gpu_model.py -s
grep -E "^depth|config1|floor" gpurun_out/model.log
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
NCU="ncu --clock-control none --profile-from-start off"
BENCH="python bench.py --profile --warmup 1 --new-tokens 64"
timeout 1500 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r01b.csv $BENCH > gpurun_out/prof_launch.log 2>&1; echo "launchlist exit $?"
for spec in "gemm_vit:gemm_tn_kernel:1:4" "gemm_prefill:gemm_tn_kernel:95:4" "gemm_decode:gemm_tn_kernel:232:4" "flash_vit:flash_fwd_kernel:1:1" "flash_prefill:flash_fwd_kernel:24:1" "decode_attn:decode_attn_kernel:40:2"; do
  IFS=: read name kn skip cnt <<< "$spec"
  timeout 900 $NCU --set full --import-source on -k regex:$kn -s $skip -c $cnt -f -o gpurun_out/r01_$name $BENCH > gpurun_out/prof_$name.log 2>&1; echo "$name exit $?"
done
ls -la gpurun_out/*.ncu-rep
