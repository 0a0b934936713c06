#!/bin/bash
# Runs every tc_harness case in its own process (a device trap kills only that case).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/harness_gpu.txt 2>&1
n=$(build/tc_harness list)
: > gpurun_out/tc_harness.log
for i in $(seq 0 $((n-1))); do
  timeout 120 build/tc_harness $i >> gpurun_out/tc_harness.log 2>&1
  rc=$?
  if [ $rc -ne 0 ]; then echo "   case $i exit code $rc" >> gpurun_out/tc_harness.log; fi
done
cat gpurun_out/tc_harness.log
