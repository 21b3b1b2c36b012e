#!/bin/bash
# 2-GPU check of the data-parallel test (gpurun --gpus 2)
mkdir -p gpurun_out
tag=${1:-r02q}
timeout 900 python -m pytest tests/test_gpu_serving.py -m gpu -q -s -p no:cacheprovider --timeout 800 > gpurun_out/pytest_2gpu_${tag}.log 2>&1; echo "serving tests (2 GPUs) exit $?"; tail -n 3 gpurun_out/pytest_2gpu_${tag}.log
