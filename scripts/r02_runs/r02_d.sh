#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/pytest_d.log 2>&1
cat gpurun_out/pytest_d.log | cut -c1-1500
for u in 2 4; do
MCN_BN_BWD_UNROLL=$u timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02d_u$u.json 2> gpurun_out/bench_r02d_u$u.err > gpurun_out/bench_r02d_u$u.json
cat gpurun_out/bench_r02d_u$u.json | cut -c1-200
tail -2 gpurun_out/bench_r02d_u$u.err
done
