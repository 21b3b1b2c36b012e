#!/bin/bash
# Round-2 check on one B200: exact-mode parity tests, whole GPU suite, smoke, racecheck (analysis report), short bench line.
mkdir -p gpurun_out
tag=${1:-r02}
timeout 1500 python -m pytest tests/test_gpu_exact.py -m gpu -q -s -p no:cacheprovider --timeout 1200 > gpurun_out/pytest_exact_${tag}.log 2>&1; echo "exact exit $?"; tail -n 5 gpurun_out/pytest_exact_${tag}.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_${tag}.log
if [ "$2" != "nosuite" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 900 --deselect tests/test_gpu_exact.py > gpurun_out/pytest_gpu_${tag}.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu_${tag}.log
fi
if [ "$3" == "race" ]; then
  SMALL='gemm_plain or gemm_epilogues or swiglu_pairs or flash_attention_tc or decode_attention or layernorm or rope_kv or swiglu_splice'
  timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 400 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "$SMALL" > gpurun_out/sanitizer_${tag}_racecheck_analysis.log 2>&1; echo "racecheck exit $?"
  grep -E "Race reported|RACECHECK SUMMARY" gpurun_out/sanitizer_${tag}_racecheck_analysis.log | sed -E 's/\+0x[0-9a-f]+//g' | sort | uniq -c | sort -rn | head -40
fi
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench exit $?"; cut -c1-3000 gpurun_out/bench_${tag}.json
