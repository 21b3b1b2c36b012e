#!/bin/bash
# tcgen05 flash attention bring-up: operand-form probes, parity tests (one process per case group), kernel timing
mkdir -p gpurun_out
P=tools/_build/probe_umma
[ "$1" == "probe" ] && {
  echo "== probes"
  for args in "0 64 16384 1024 2048 0" "0 128 16384 1024 2048 0" "1 64 16384 1024 2048 0" "1 128 16384 1024 2048 0" \
              "0 128 1024 16384 2048 0" "0 64 1024 16384 2048 0" "1 64 16384 1024 2048 1" "0 64 16384 1024 32 0" "0 64 1024 1024 2048 0"; do
    timeout 60 $P $args; echo "  exit $?"
  done
} > gpurun_out/probe.log 2>&1
cat gpurun_out/probe.log
for k in "tc and 64-" "tc and 128-" "tc_matches"; do
  timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 tests/test_gpu_ops.py -k "$k" -x > "gpurun_out/fa_${k// /_}.log" 2>&1
  echo "tests [$k] exit $?"; tail -n 12 "gpurun_out/fa_${k// /_}.log" | cut -c1-300
done
timeout 300 python tools/fa_bench.py vit 2>&1 | tail -n 8
timeout 300 python tools/fa_bench.py prefill 2>&1 | tail -n 8
