#!/bin/bash
# Re-entry check of HEAD on one B200: whole GPU suite, smoke, default bench line, flash-attention and pair-GEMM stand-alone timings.
mkdir -p gpurun_out
tag=${1:-r02z}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_${tag}.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu_${tag}.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_${tag}.log
timeout 1200 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cut -c1-900 gpurun_out/bench_${tag}.json
timeout 300 python tools/fa_bench.py all 2>&1 | grep -v "^$" | tee gpurun_out/fa_${tag}.log
timeout 600 python tools/pair_sweep.py cublas prefill 2>&1 | tail -n 30 | tee gpurun_out/pair_cublas_prefill_${tag}.log
timeout 600 python tools/pair_sweep.py cublas vit 2>&1 | tail -n 30 | tee gpurun_out/pair_cublas_vit_${tag}.log
