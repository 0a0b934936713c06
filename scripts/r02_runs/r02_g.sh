#!/bin/bash
mkdir -p gpurun_out
(time timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/pytest_g.log 2>&1
cat gpurun_out/pytest_g.log | cut -c1-1200
for c in r50 r50_fp32_b32 effnet_b0 deeplab_r50_512 dcgan_64; do
timeout 600 python bench.py --config $c --no-cpu-baseline --steps 8 --profile-json gpurun_out/prof_r02g_$c.json 2> gpurun_out/bench_r02g_$c.err > gpurun_out/bench_r02g_$c.json
echo "== $c rc=$?"; cat gpurun_out/bench_r02g_$c.json | cut -c1-160
tail -2 gpurun_out/bench_r02g_$c.err | cut -c1-300
done
