#!/bin/bash
# 2-GPU check of the final build (gpurun --gpus 2): the data-parallel serving tests (2 ranks == 1 rank) and the bench line under torchrun.
mkdir -p gpurun_out
tag=${1:-r02m3}
timeout 900 python -m pytest tests/test_gpu_serving.py -m gpu -q -p no:cacheprovider --timeout 800 > gpurun_out/pytest_2gpu_${tag}.log 2>&1; echo "serving tests (2 GPUs) exit $?"; tail -n 3 gpurun_out/pytest_2gpu_${tag}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_${tag}.json 2> gpurun_out/bench_2gpu_${tag}.err; echo "bench exit $?"; tail -n 1 gpurun_out/bench_2gpu_${tag}.json | cut -c1-600
