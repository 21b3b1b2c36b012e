#!/bin/bash
# Verification of HEAD on one B200 (GPU suite, smoke, default bench line) + tensor-pipe activity of gemm_pair_kernel on all four prefill shapes.
mkdir -p gpurun_out
tag=${1:-r02k}
bash scripts/gpu_r02_final2.sh ${tag}
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,lts__t_bytes.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct"
timeout 900 ncu --clock-control none --metrics $M -k regex:'gemm_pair_kernel' -c 24 --csv --log-file gpurun_out/pair_ncu_${tag}.csv python tools/pair_sweep.py cublas prefill > gpurun_out/pair_ncu_${tag}.log 2>&1; echo "ncu exit $?"
timeout 900 ncu --clock-control none --metrics $M -k regex:'gemm_pair_kernel' -c 24 --csv --log-file gpurun_out/pair_ncu_vit_${tag}.csv python tools/pair_sweep.py cublas vit > gpurun_out/pair_ncu_vit_${tag}.log 2>&1; echo "ncu exit $?"
python - <<PY
import csv
for f in ("gpurun_out/pair_ncu_${tag}.csv", "gpurun_out/pair_ncu_vit_${tag}.csv"):
    rows=[r for r in csv.DictReader(l for l in open(f) if not l.startswith("=="))]
    k={}
    for r in rows:
        k.setdefault((int(r["ID"]), r["Kernel Name"][:32]), {})[r["Metric Name"]]=r["Metric Value"]
    for (i,n),m in sorted(k.items()):
        if i % 6 == 0:
            print(i, n, "us", m.get("gpu__time_duration.sum"), "tensor%", m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), "cycles", m.get("sm__cycles_elapsed.max"), "lts_bytes", m.get("lts__t_bytes.sum"), "dram_rd", m.get("dram__bytes_read.sum"))
PY
