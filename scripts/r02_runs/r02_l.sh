#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_l.log
grep -n "Error\|assert\|FAILED\|passed\|failed" gpurun_out/pytest_l.log | cut -c1-1200 | head -60
for f in 1 0; do
MCN_FUSE_BN_POOL=$f timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02l_f$f.json 2> gpurun_out/bench_r02l_f$f.err > gpurun_out/bench_r02l_f$f.json
cat gpurun_out/bench_r02l_f$f.json | cut -c1-140
done
