#!/bin/bash
mkdir -p gpurun_out
n=$(build/tc_harness list)
for bo in 0 1; do
  echo "=== MCN_HALO_BASE_OFFSET=$bo" >> gpurun_out/halo.log
  for i in $(seq 25 37); do
    MCN_HALO_BASE_OFFSET=$bo timeout 120 build/tc_harness $i >> gpurun_out/halo.log 2>&1 || echo "   case $i rc $?" >> gpurun_out/halo.log
  done
done
cat gpurun_out/halo.log
