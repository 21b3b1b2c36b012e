#!/bin/bash
# ncu --set full of the CTA-pair GEMM at the bench shapes: two ViT launches and two prefill launches of one bench step
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off --set full --import-source on -k regex:gemm_pair_kernel"
BENCH="python bench.py --profile --warmup 1 --new-tokens 5"
timeout 600 $NCU -s 10 -c 2 -f -o gpurun_out/r01_gemm_pair_vit $BENCH > gpurun_out/ncu_gemm_pair_vit.log 2>&1; echo "vit exit $?"
timeout 600 $NCU -s 101 -c 4 -f -o gpurun_out/r01_gemm_pair_prefill $BENCH > gpurun_out/ncu_gemm_pair_prefill.log 2>&1; echo "prefill exit $?"
ls -la gpurun_out/*.ncu-rep
