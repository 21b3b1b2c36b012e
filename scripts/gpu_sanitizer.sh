#!/bin/bash
# compute-sanitizer over the kernel-level parity tests: memcheck on all of tests/test_gpu_ops.py,
# racecheck + synccheck on the tests that exercise the hand-rolled mbarrier / TMEM protocols at small shapes.
# Usage: scripts/gpu_sanitizer.sh [tag]   -> gpurun_out/sanitizer_<tag>_{memcheck,racecheck,synccheck}.log
tag=${1:-r02}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest tests/test_gpu_ops.py -m gpu -x -q -p no:cacheprovider"
SMALL='gemm_plain or gemm_epilogues or swiglu_pairs or flash_attention_tc or decode_attention or layernorm or rope_kv or swiglu_splice'
timeout 1500 $CS --tool memcheck --error-exitcode 9 --print-limit 50 $PY > gpurun_out/sanitizer_${tag}_memcheck.log 2>&1; echo "memcheck exit $?"
tail -n 6 gpurun_out/sanitizer_${tag}_memcheck.log
timeout 1500 $CS --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 50 $PY -k "$SMALL" > gpurun_out/sanitizer_${tag}_racecheck.log 2>&1; echo "racecheck exit $?"
tail -n 6 gpurun_out/sanitizer_${tag}_racecheck.log
timeout 900 $CS --tool synccheck --error-exitcode 9 --print-limit 50 $PY -k "$SMALL" > gpurun_out/sanitizer_${tag}_synccheck.log 2>&1; echo "synccheck exit $?"
tail -n 6 gpurun_out/sanitizer_${tag}_synccheck.log
