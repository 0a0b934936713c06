#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_u1.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_u1.log | cut -c1-500 | head -20
for p in 1 0; do
MCN_BN_PIPE=$p timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02u_p$p.json 2> gpurun_out/bench_r02u_p$p.err > gpurun_out/bench_r02u_p$p.json
grep "timed region\|end-to-end" gpurun_out/bench_r02u_p$p.err; tail -2 gpurun_out/bench_r02u_p$p.err
done
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_u.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_u.log | cut -c1-700 | head -40
