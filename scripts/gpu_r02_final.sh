#!/bin/bash
# Round-2 final check on one B200: whole GPU suite, smoke, default bench line (with cpu_baseline), reference arm (3 steps), launch list,
# ncu --set full of the decode kernels that changed this round.
mkdir -p gpurun_out
tag=${1:-r02z}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_${tag}.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu_${tag}.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_${tag}.log
timeout 1200 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cut -c1-1200 gpurun_out/bench_${tag}.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "ref exit $?"; cut -c1-900 gpurun_out/bench_ref_${tag}.json
timeout 900 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --profile --warmup 1 --new-tokens 5 > gpurun_out/prof_launch_${tag}.log 2>&1; echo "launchlist exit $?"
NCU="ncu --clock-control none --profile-from-start off --set full --import-source on"
BENCH="python bench.py --profile --warmup 1 --new-tokens 5"
timeout 600 $NCU -k regex:decode_attn_mma_kernel -s 40 -c 1 -f -o gpurun_out/${tag}_decode_attn_mma $BENCH > gpurun_out/ncu_${tag}_attn.log 2>&1; echo "ncu attn exit $?"
timeout 600 $NCU -k regex:gemm_tn_kernel -s 200 -c 4 -f -o gpurun_out/${tag}_gemm_decode $BENCH > gpurun_out/ncu_${tag}_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 600 $NCU -k "regex:reduce_rope_kv_write_kernel|reduce_swiglu_kernel|reduce_residual_rmsnorm_kernel" -s 60 -c 3 -f -o gpurun_out/${tag}_decode_glue $BENCH > gpurun_out/ncu_${tag}_glue.log 2>&1; echo "ncu glue exit $?"
ls -la gpurun_out/*.ncu-rep | tail -5
