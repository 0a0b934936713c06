#!/bin/bash
# One GPU visit: harness cases given as arguments, then the GPU test-suite, then the bench with
# the per-launch profile.  If a halo-wgrad harness case fails the rest runs with MCN_WGRAD_HALO=0.
mkdir -p gpurun_out
log=gpurun_out/harness_sel.log
: > $log
for i in "$@"; do
  timeout 60 build/tc_harness $i >> $log 2>&1 || echo "   case $i rc $?" >> $log
done
cat $log
if grep -q "FAIL\|ERROR\|FAULT\|rc " $log; then export MCN_WGRAD_HALO=0; echo "== halo wgrad disabled"; fi
(time timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/pytest.log 2>&1
cat gpurun_out/pytest.log
timeout 120 python bench.py --no-cpu-baseline --profile-json gpurun_out/prof_new.json > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err
cat gpurun_out/bench_new.json; tail -4 gpurun_out/bench_new.err
