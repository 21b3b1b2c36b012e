#!/bin/bash
# tests + bench (+ optional reference-arm timing) on one GPU
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 python -m pytest -q -m gpu -p no:cacheprovider --timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?"; tail -n 3 gpurun_out/$name.log; }
run ops tests/test_gpu_ops.py
run model tests/test_gpu_model.py -s
grep -E "^depth|config1 step|floor|verified" gpurun_out/model.log | cut -c1-260
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "ref" ]; then
  nproc; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"
  ( time timeout 1200 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -n 4 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
fi
