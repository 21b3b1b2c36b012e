#!/bin/bash
# Round-2 evidence for the kernels changed late in the round (flash attention issuer, pair GEMM epilogue / 512-row tiles):
# read-only HBM ceilings + decode-attention streaming skeleton, flash timing decomposition, ncu --set full of the flash
# kernels, of the pair GEMM next to cuBLAS on the same operands, and the launch list of one step.
mkdir -p gpurun_out
tag=${1:-r02e}
timeout 300 tools/_build/hbm_read_peak > gpurun_out/hbm_read_peak_${tag}.log 2>&1; echo "hbm exit $?"; cat gpurun_out/hbm_read_peak_${tag}.log
for v in "" 1; do
  TEO_DEC_DBG_SKIP=$v timeout 300 python tools/dec_attn_bench.py 2>&1 | tail -n 1 | sed "s/^/skip=[$v] /" | tee -a gpurun_out/dec_attn_${tag}.log
done
bash scripts/gpu_r02_fa_dbg.sh ${tag}
for W in vit prefill; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:flash_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_flash_tc_$W python tools/fa_bench.py $W > gpurun_out/ncu_fa_$W.log 2>&1; echo "ncu fa $W exit $?"
done
timeout 900 ncu --set full --clock-control none -k regex:'gemm_pair_kernel|nvjet|cutlass|sm100|gemm' -s 2 -c 4 -f -o gpurun_out/r02_pair_vs_cublas python tools/pair_sweep.py cublas prefill > gpurun_out/ncu_pair_cublas.log 2>&1; echo "ncu pair exit $?"
bash scripts/gpu_launchlist.sh 5 ${tag}
ls -la gpurun_out/*.ncu-rep
