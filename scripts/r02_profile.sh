#!/bin/bash
# Round-2 evidence run: bench lines of all five BASELINE configs, role timing, ncu launch list + full set.
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02/gpu.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/r02/kernel_classes_r50.json > gpurun_out/r02/bench_r50.json 2> gpurun_out/r02/bench_r50.err
cut -c1-160 gpurun_out/r02/bench_r50.json
for c in r50_fp32_b32 effnet_b0 deeplab_r50_512 dcgan_64; do
timeout 600 python bench.py --config $c --no-cpu-baseline --steps 10 --profile-json gpurun_out/r02/kernel_classes_$c.json > gpurun_out/r02/bench_$c.json 2> gpurun_out/r02/bench_$c.err
cut -c1-160 gpurun_out/r02/bench_$c.json
done
MCN_LIB=$PWD/myconvnet_b200/libmcn_timing.so timeout 600 python scripts/role_timing.py > gpurun_out/r02/role_timing.txt 2> gpurun_out/r02/role_timing.err
bash scripts/ncu_profile.sh
