#!/bin/bash
# round 2, first GPU visit: determinism of smoke(), the whole GPU suite (no -x), a short bench
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 > gpurun_out/smoke_$i.log; done
cat gpurun_out/smoke_1.log gpurun_out/smoke_2.log
(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest.log 2>&1
cat gpurun_out/pytest.log | cut -c1-400
timeout 300 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02a.json 2> gpurun_out/bench_r02a.err > gpurun_out/bench_r02a.json
cat gpurun_out/bench_r02a.json | cut -c1-1500
tail -5 gpurun_out/bench_r02a.err
