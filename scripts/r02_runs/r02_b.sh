#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02b.json 2> gpurun_out/bench_r02b.err > gpurun_out/bench_r02b.json
cat gpurun_out/bench_r02b.json | cut -c1-400
tail -3 gpurun_out/bench_r02b.err
timeout 900 python scripts/measure_parity.py 32 > gpurun_out/parity_config1.txt 2>&1
grep -v "shape\|drop rate" gpurun_out/parity_config1.txt | cut -c1-900
