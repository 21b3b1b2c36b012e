#!/bin/bash
# Evidence for the final build of round 2: compute-sanitizer over the GEMM tests (the staged epilogue and the stream-K fix-up changed),
# ncu --set full of gemm_pair_kernel<2> / <1> at a prefill and a ViT shape, launch list of one step.
mkdir -p gpurun_out
tag=${1:-r02fin}
CS=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest tests/test_gpu_ops.py -m gpu -x -q -p no:cacheprovider"
K='gemm or swiglu_pairs or layernorm'
timeout 900 $CS --tool memcheck --error-exitcode 9 --print-limit 20 $PY -k "$K" > gpurun_out/sanitizer_${tag}_memcheck.log 2>&1; echo "memcheck exit $?"; tail -n 4 gpurun_out/sanitizer_${tag}_memcheck.log
timeout 900 $CS --tool synccheck --error-exitcode 9 --print-limit 20 $PY -k "$K" > gpurun_out/sanitizer_${tag}_synccheck.log 2>&1; echo "synccheck exit $?"; tail -n 4 gpurun_out/sanitizer_${tag}_synccheck.log
timeout 900 $CS --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 20 $PY -k "gemm_plain or gemm_epilogues or swiglu_pairs" > gpurun_out/sanitizer_${tag}_racecheck.log 2>&1; echo "racecheck exit $?"; grep -c "Race reported" gpurun_out/sanitizer_${tag}_racecheck.log; grep "Race reported" gpurun_out/sanitizer_${tag}_racecheck.log | sed 's/+0x.*//' | sort | uniq -c | head; tail -n 3 gpurun_out/sanitizer_${tag}_racecheck.log
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:gemm_pair_kernel -f -o gpurun_out/r02_gemm_pair_prefill python tools/pair_sweep.py one 0 0 -1 -1 prefill > gpurun_out/ncu_pair_prefill_${tag}.log 2>&1; echo "ncu prefill exit $?"
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:gemm_pair_kernel -f -o gpurun_out/r02_gemm_pair_vit python tools/pair_sweep.py one 0 0 -1 -1 vit > gpurun_out/ncu_pair_vit_${tag}.log 2>&1; echo "ncu vit exit $?"
bash scripts/gpu_launchlist.sh 5 ${tag}
ls -la gpurun_out/*.ncu-rep
