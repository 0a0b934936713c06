#!/bin/bash
mkdir -p gpurun_out
MCN_LIB=$PWD/myconvnet_b200/libmcn_timing.so timeout 600 python scripts/role_timing.py > gpurun_out/role_timing.txt 2> gpurun_out/role_timing.err
cat gpurun_out/role_timing.txt; tail -5 gpurun_out/role_timing.err
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short -x 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_o.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_o.log | cut -c1-700 | head -40
timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02o.json 2> gpurun_out/bench_r02o.err > gpurun_out/bench_r02o.json
cut -c1-200 gpurun_out/bench_r02o.json
