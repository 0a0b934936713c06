#!/bin/bash
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -60) > gpurun_out/pytest_c.log 2>&1
cat gpurun_out/pytest_c.log | cut -c1-1200
timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02c.json 2> gpurun_out/bench_r02c.err > gpurun_out/bench_r02c.json
cat gpurun_out/bench_r02c.json | cut -c1-300
tail -2 gpurun_out/bench_r02c.err
