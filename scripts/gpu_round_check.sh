#!/bin/bash
# Round check on one B200: host facts, GPU parity tests, smoke, default bench line, reference arm (1 step), short launch list.
mkdir -p gpurun_out
{ nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)|^CPU\(s\)"; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader; } > gpurun_out/host.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_r01d.json
if [ "$1" != "nolist" ]; then
  timeout 900 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r01d.csv python bench.py --profile --warmup 1 --new-tokens 5 > gpurun_out/prof_launch_r01d.log 2>&1; echo "launchlist exit $?"
fi
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r01d.json 2> gpurun_out/bench_ref_r01d.err; echo "ref exit $?"; cut -c1-400 gpurun_out/bench_ref_r01d.json
