#!/bin/bash
# Round-2 check 3: whole GPU suite, decode-GEMM skew diagnostic (per-SM stream times), cuBLAS comparison of the tiled GEMM.
mkdir -p gpurun_out
tag=${1:-r02g}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_${tag}.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu_${tag}.log
timeout 400 python tools/dec_gemm_skew.py 32 > gpurun_out/dec_gemm_skew_${tag}.log 2>&1; echo "skew exit $?"; cat gpurun_out/dec_gemm_skew_${tag}.log
timeout 300 python tools/pair_sweep.py cublas vit > gpurun_out/pair_vs_cublas_${tag}.log 2>&1; echo "cublas vit exit $?"
timeout 300 python tools/pair_sweep.py cublas prefill >> gpurun_out/pair_vs_cublas_${tag}.log 2>&1; echo "cublas prefill exit $?"; cat gpurun_out/pair_vs_cublas_${tag}.log
