#!/bin/bash
# Per-tile timeline of the pair GEMM (clock64 stamps, trace build): 512-row tiles on the prefill shapes, both tilings at K = 1024.
mkdir -p gpurun_out
tag=${1:-r02o}
export TEO_LIB_PATH=teochat_b200/lib/variants/pairtrace.so
timeout 300 python tools/pair_trace.py prefill 2>&1 | tee gpurun_out/pair_trace_prefill_${tag}.log | cut -c1-420
TEO_PAIR_MT=1024 timeout 300 python tools/pair_trace.py vit qkv 2>&1 | tee gpurun_out/pair_trace_vit_mt2_${tag}.log | cut -c1-420
timeout 300 python tools/pair_trace.py vit qkv 2>&1 | tee gpurun_out/pair_trace_vit_mt1_${tag}.log | cut -c1-420
