#!/bin/bash
# Flash-attention kernel A/B on one box: variant lib (TEO_LIB_PATH=$1, default variants/fa_old.so) against the tree's build.
mkdir -p gpurun_out
tag=${2:-r02v}
V=${1:-teochat_b200/lib/variants/fa_old.so}
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "flash or attention" 2>&1 | tail -3
for i in 1 2 3; do
  TEO_LIB_PATH=$V timeout 300 python tools/fa_bench.py all 2>&1 | grep -v "^$" | sed "s/^/old $i: /" | tee -a gpurun_out/fa_${tag}.log
  timeout 300 python tools/fa_bench.py all 2>&1 | grep -v "^$" | sed "s/^/new $i: /" | tee -a gpurun_out/fa_${tag}.log
done
