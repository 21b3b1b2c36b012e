#!/bin/bash
# ncu --set full of the tcgen05 flash kernels at the bench shapes (args: vit|prefill)
mkdir -p gpurun_out
W=${1:-vit}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:flash_tc_kernel -s 3 -c 1 -f -o gpurun_out/r01_flash_tc_$W python tools/fa_bench.py $W > gpurun_out/ncu_fa_$W.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_fa_$W.log; ls -la gpurun_out/*.ncu-rep
