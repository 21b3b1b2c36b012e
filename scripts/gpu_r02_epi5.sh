#!/bin/bash
# Tensor-pipe activity of the 512-row pair tiles (TEO_PAIR_MT=4096) and of the cuBLAS kernels on the same operands, all four prefill shapes.
mkdir -p gpurun_out
tag=${1:-r02h}
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,lts__t_bytes.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
TEO_PAIR_MT=4096 timeout 900 ncu --clock-control none --metrics $M -k regex:'gemm_pair_kernel|nvjet' -c 48 --csv --log-file gpurun_out/mt2_ncu_${tag}.csv python tools/pair_sweep.py cublas prefill > gpurun_out/mt2_ncu_${tag}.log 2>&1; echo "ncu exit $?"
python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/mt2_ncu_${tag}.csv") if not l.startswith("=="))]
from collections import OrderedDict
k=OrderedDict()
for r in rows:
    k.setdefault((r["ID"], r["Kernel Name"][:40]), {})[r["Metric Name"]]=r["Metric Value"]
seen=set()
for (i,n),m in k.items():
    key=(n, m.get("sm__cycles_elapsed.max","")[:3])
    print(i, n, m.get("gpu__time_duration.sum"), "tensor%", m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), "cycles", m.get("sm__cycles_elapsed.max"), "lts_bytes", m.get("lts__t_bytes.sum"), "dram_rd", m.get("dram__bytes_read.sum"), "hit", m.get("lts__t_sector_hit_rate.pct"), "smem_wavefronts", m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"))
PY
