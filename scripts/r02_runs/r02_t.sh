#!/bin/bash
mkdir -p gpurun_out
for p in 1 0 1 0; do
MCN_PDL=$p timeout 600 python bench.py --no-cpu-baseline --steps 10 2> gpurun_out/bench_r02t_p$p.err > gpurun_out/bench_r02t_p$p.json
grep "timed region\|end-to-end" gpurun_out/bench_r02t_p$p.err
done
