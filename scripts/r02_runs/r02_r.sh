#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "bn or batch" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_r1.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_r1.log | cut -c1-500 | head -20
for u in 4 2; do
MCN_BN_NOY_UNROLL=$u timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02r_u$u.json 2> gpurun_out/bench_r02r_u$u.err > gpurun_out/bench_r02r_u$u.json
cut -c1-200 gpurun_out/bench_r02r_u$u.json
done
