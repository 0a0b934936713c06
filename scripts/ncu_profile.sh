#!/bin/bash
# Round profile: launch list of one bench run + one --set full capture of the top kernels.
# Each ncu pass only after the identical plain command exited 0 (B200_PROFILING.md).
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 480 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_conv_kernel|wgrad_kernel|bn_bwd_apply_kernel|bn_bwd_reduce_kernel|bn_apply_kernel" \
    -s 230 -c 36 -o gpurun_out/prof_full $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/
