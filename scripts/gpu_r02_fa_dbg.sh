#!/bin/bash
# Timing decomposition of the tcgen05 flash kernels: debug variants (lib/variants/fa_dbg{1,2,3}.so, -DTEO_FA_DBG=n) against the tree's build.
mkdir -p gpurun_out
tag=${1:-r02w}
for v in "" fa_dbg1 fa_dbg2 fa_dbg3 ""; do
  if [ -z "$v" ]; then pre="default"; unset TEO_LIB_PATH; else pre=$v; export TEO_LIB_PATH=teochat_b200/lib/variants/$v.so; fi
  timeout 300 python tools/fa_bench.py all 2>&1 | grep "tcgen05" | sed "s/^/$pre: /" | tee -a gpurun_out/fa_dbg_${tag}.log
done
