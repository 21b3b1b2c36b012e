#!/bin/bash
# first GPU validation: each group in its own process so a trap in one kernel cannot poison the others
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run misc tests/test_gpu_ops.py -k "hash or patchify or layernorm or assemble or swiglu or rope"
run gemm tests/test_gpu_ops.py -k "gemm"
run flash tests/test_gpu_ops.py -k "flash"
run decode tests/test_gpu_ops.py -k "decode"
run model tests/test_gpu_model.py -s
cat gpurun_out/summary.txt
for f in misc gemm flash decode model; do echo "=== $f"; tail -n 25 gpurun_out/$f.log; done
