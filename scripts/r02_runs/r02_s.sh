#!/bin/bash
mkdir -p gpurun_out
for p in 1 0; do
MCN_PDL=$p timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02s_p$p.json 2> gpurun_out/bench_r02s_p$p.err > gpurun_out/bench_r02s_p$p.json
cut -c1-200 gpurun_out/bench_r02s_p$p.json; tail -3 gpurun_out/bench_r02s_p$p.err
done
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short -x 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_s.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_s.log | cut -c1-700 | head -40
