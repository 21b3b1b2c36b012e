#!/bin/bash
# Ring depth of the 512-row pair tiles: the tree's build (two staging buffers per epilogue warp: 3 stages of 48 KiB) against the
# TEO_EPI_BUFS=1 variant (one buffer: 4 stages, no residual prefetch), both with 512-row tiles for K >= 4096, against cuBLAS.
mkdir -p gpurun_out
tag=${1:-r02g}
for i in 1 2; do
  for v in "" teochat_b200/lib/variants/epi1.so; do
    TEO_LIB_PATH=$v TEO_PAIR_MT=4096 timeout 600 python tools/pair_sweep.py cublas prefill 2>&1 | grep "ours/cuBLAS" | sed "s|^|lib=[$v] |" | tee -a gpurun_out/epi4_${tag}.log
  done
done
