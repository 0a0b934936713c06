#!/bin/bash
mkdir -p gpurun_out
timeout 600 bash scripts/run_tc_harness.sh > /dev/null 2>&1
grep -c PASS gpurun_out/tc_harness.log; grep -v PASS gpurun_out/tc_harness.log | head -40
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_p1.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_p1.log | cut -c1-500 | head -20
MCN_LIB=$PWD/myconvnet_b200/libmcn_timing.so timeout 600 python scripts/role_timing.py > gpurun_out/role_timing_p.txt 2> gpurun_out/role_timing_p.err
cat gpurun_out/role_timing_p.txt; tail -5 gpurun_out/role_timing_p.err
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_p.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_p.log | cut -c1-700 | head -40
for t in 1 0; do
MCN_TMA_STORE=$t timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02p_t$t.json 2> gpurun_out/bench_r02p_t$t.err > gpurun_out/bench_r02p_t$t.json
cut -c1-200 gpurun_out/bench_r02p_t$t.json
done
