#!/bin/bash
# launch list of one profiled step (args: new_tokens tag)
mkdir -p gpurun_out
NT=${1:-16}; TAG=${2:-x}
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --profile --warmup 1 --new-tokens $NT > gpurun_out/prof_launch_$TAG.log 2>&1; echo "launchlist exit $?"
python tools/launch_summary.py gpurun_out/launches_$TAG.csv | tail -n 22
