#!/bin/bash
# Final verification of HEAD on one B200: whole GPU suite, smoke, default bench line.
mkdir -p gpurun_out
tag=${1:-r02y}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_${tag}.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu_${tag}.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_${tag}.log
timeout 1200 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cut -c1-700 gpurun_out/bench_${tag}.json
