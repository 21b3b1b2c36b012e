#!/bin/bash
# 2-GPU check (gpurun --gpus 2): 2 ranks == 1 rank (NCCL gather of ids), weak-scaling bench line with per-rank phases, configs[4] under torchrun.
mkdir -p gpurun_out
tag=${1:-r02n}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_serving.py -m gpu -q -s -p no:cacheprovider --timeout 800 -k "two_ranks" > gpurun_out/pytest_2gpu_${tag}.log 2>&1; echo "2-rank test exit $?"; tail -n 3 gpurun_out/pytest_2gpu_${tag}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/bench_n2_${tag}.json 2> gpurun_out/bench_n2_${tag}.err; echo "bench N=2 exit $?"; cut -c1-1500 gpurun_out/bench_n2_${tag}.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 2 --warmup 2 --config 4 > gpurun_out/bench_n2_c4_${tag}.json 2> gpurun_out/bench_n2_c4_${tag}.err; echo "bench N=2 configs[4] exit $?"; cut -c1-600 gpurun_out/bench_n2_c4_${tag}.json
