#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/harness_epi.log; : > $log
for i in $(seq 0 33) 46 47 48 49 50 51 52 58 59 60; do timeout 120 build/tc_harness $i >> $log 2>&1 || echo "   case $i rc $?" >> $log; done
grep -c PASS $log; grep -E "FAIL|ERROR|FAULT|rc " $log | head
for i in 61 62 63 64 38 39 34 37; do timeout 60 build/tc_harness $i 2>&1; done
(timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6) | cut -c1-300
MCN_WEIGHT_STATIONARY=2 timeout 300 python -m pytest tests/test_gpu_models.py tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -3 | cut -c1-300
timeout 120 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_epi.json 2> gpurun_out/bench_epi.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
