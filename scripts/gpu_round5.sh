#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/harness_ws.log; : > $log
for i in 58 59 60; do timeout 120 build/tc_harness $i >> $log 2>&1 || echo "   case $i rc $?" >> $log; done
for ws in 1 0; do echo "== MCN_WEIGHT_STATIONARY=$ws" >> $log; for i in 61 62 63 64 38 39; do MCN_WEIGHT_STATIONARY=$ws timeout 60 build/tc_harness $i >> $log 2>&1; done; done
cat $log
echo "== pytest with forced weight-stationary"
MCN_WEIGHT_STATIONARY=2 timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
for ws in 1 0; do
  MCN_WEIGHT_STATIONARY=$ws timeout 120 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_ws$ws.json 2> gpurun_out/bench_ws$ws.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('WS=$ws', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
