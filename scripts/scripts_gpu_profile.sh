#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest -q -m gpu -p no:cacheprovider --timeout 900 tests/test_gpu_model.py -k "config1 or full_width" -s > gpurun_out/model_full.log 2>&1; echo "model_full exit $?"
grep -E "depth|config1|oracle repro|passed|failed" gpurun_out/model_full.log | head -20
NCU="ncu --clock-control none --profile-from-start off"
BENCH="python bench.py --profile --warmup 1 --new-tokens 64"
timeout 1500 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r01.csv $BENCH > gpurun_out/prof_launch.log 2>&1; echo "launchlist exit $?"
tail -n 2 gpurun_out/prof_launch.log
for spec in "decode_attn:regex:decode_attn_kernel:40" "gemm256:regex:gemm_tn_kernelILi256:30" "flash128:regex:flash_fwd_kernelILi128:4" "flash64:regex:flash_fwd_kernelILi64:4" "gemm32:regex:gemm_tn_kernelILi32:40"; do
  IFS=: read name r1 r2 skip <<< "$spec"
  timeout 900 $NCU --set full --import-source on -k $r1:$r2 -s $skip -c 2 -f -o gpurun_out/r01_$name $BENCH > gpurun_out/prof_$name.log 2>&1; echo "$name exit $?"
done
ls -la gpurun_out/*.ncu-rep
