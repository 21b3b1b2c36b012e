#!/bin/bash
# Round-2 check: whole GPU suite (not -x), smoke, cuBLAS comparison of the tiled GEMM at the ViT / prefill shapes.
mkdir -p gpurun_out
tag=${1:-r02f}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_${tag}.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu_${tag}.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_${tag}.log
timeout 300 python tools/pair_sweep.py cublas vit > gpurun_out/pair_vs_cublas_${tag}.log 2>&1; echo "cublas vit exit $?"
timeout 300 python tools/pair_sweep.py cublas prefill >> gpurun_out/pair_vs_cublas_${tag}.log 2>&1; echo "cublas prefill exit $?"; cat gpurun_out/pair_vs_cublas_${tag}.log
