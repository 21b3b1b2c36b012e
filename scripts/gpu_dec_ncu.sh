#!/bin/bash
# ncu --set full of the decode attention kernel at the bench shape
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_attn -s 4 -c 1 -f -o gpurun_out/r01_decode_attn_mma python tools/dec_attn_bench.py > gpurun_out/ncu_dec.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_dec.log
