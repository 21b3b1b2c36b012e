#!/bin/bash
# Tensor-pipe activity of gemm_pair_kernel, one isolated launch per shape (tools/pair_sweep.py one): prefill shapes (512-row tiles), ViT shapes with
# 256-row (default) and 512-row tiles (TEO_PAIR_MT=1024).
mkdir -p gpurun_out
tag=${1:-r02m}
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,lts__t_bytes.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct"
run() { # name env which
  env $2 timeout 600 ncu --clock-control none --profile-from-start off --metrics $M -k regex:'gemm_pair_kernel' --csv --log-file gpurun_out/one_$1_${tag}.csv python tools/pair_sweep.py one 0 0 -1 -1 $3 > gpurun_out/one_$1_${tag}.log 2>&1
  python - "$1" "gpurun_out/one_$1_${tag}.csv" <<PY
import csv,sys
rows=[r for r in csv.DictReader(l for l in open(sys.argv[2]) if not l.startswith("=="))]
k={}
for r in rows: k.setdefault((int(r["ID"]), r["Kernel Name"][:28]), {})[r["Metric Name"]]=r["Metric Value"]
for (i,n),m in sorted(k.items()):
    print(sys.argv[1], i, n, "us", m.get("gpu__time_duration.sum"), "tensor%", m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), "cycles", m.get("sm__cycles_elapsed.max"), "lts", m.get("lts__t_bytes.sum"), "dram_rd", m.get("dram__bytes_read.sum"))
PY
}
run prefill_mt2 "TEO_X=1" prefill | tee -a gpurun_out/epi8_${tag}.log
run prefill_mt1 "TEO_PAIR_MT=1" prefill | tee -a gpurun_out/epi8_${tag}.log
run vit_mt1 "TEO_X=1" vit | tee -a gpurun_out/epi8_${tag}.log
run vit_mt2 "TEO_PAIR_MT=1024" vit | tee -a gpurun_out/epi8_${tag}.log
