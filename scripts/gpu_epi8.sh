#!/bin/bash
# A/B of the experimental two-epilogue-warps-per-quadrant mode (MCN_EPI_WARPS=8) on the harness.
mkdir -p gpurun_out
log=gpurun_out/harness_epi8.log; : > $log
for i in $(seq 0 24); do MCN_EPI_WARPS=8 timeout 60 build/tc_harness $i >> $log 2>&1 || echo "   case $i rc $?" >> $log; done
echo "epi8 correctness: $(grep -c PASS $log) pass"; grep -E "FAIL|ERROR|FAULT|rc " $log | head -5
for w in 4 8; do echo "== MCN_EPI_WARPS=$w"; for i in 61 62 63 64 38 39; do MCN_EPI_WARPS=$w timeout 60 build/tc_harness $i 2>&1; done; done | tee -a $log
