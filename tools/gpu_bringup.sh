#!/bin/bash
# one gpurun call: every self-test group in its own process (a trap in one group must not poison the others)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for g in "$@"; do
  echo "=== $g" | tee -a gpurun_out/selftest.log
  timeout 300 python tools/gpu_selftest.py $g >> gpurun_out/selftest.log 2>&1
  echo "exit $?" | tee -a gpurun_out/selftest.log
done
grep -E "^(PASS|FAIL|SELFTEST|PERF|===|exit)" gpurun_out/selftest.log | tail -150
