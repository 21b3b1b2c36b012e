#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 600 tests/test_gpu_model.py -k config1 -s > gpurun_out/config1.log 2>&1; echo "config1 exit $?"
tail -n 8 gpurun_out/config1.log
timeout 1500 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench exit $?"
tail -n 5 gpurun_out/bench1.err; cat gpurun_out/bench1.json
