#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_bn.py 2>&1 | grep -v "shape \[\|drop rate" | tail -40 > gpurun_out/debug_bn.txt
cat gpurun_out/debug_bn.txt
for a in 0 1; do
timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02j_a$a.json 2> gpurun_out/bench_r02j_a$a.err > gpurun_out/bench_r02j_a$a.json
cat gpurun_out/bench_r02j_a$a.json | cut -c1-140
done
