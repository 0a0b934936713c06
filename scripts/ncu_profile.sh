#!/bin/bash
# Round profile: (1) launch list of one eager bench step, (2) --set full on a handful of launches
# of the heaviest kernels, exported to CSV on the box (the .ncu-rep itself is too large to bring
# back).  Each ncu pass only after the identical plain command exited 0 (B200_PROFILING.md).
mkdir -p gpurun_out/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/ncu/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1830 -c 460 --csv \
    --log-file gpurun_out/ncu/launches.csv $CMD > gpurun_out/ncu/ncu1.log 2>&1
$CMD > gpurun_out/ncu/plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:"wgrad_kernel|gemm_conv_kernel|bn_bwd_apply_kernel|bn_bwd_reduce_kernel|bn_apply_kernel|bn_stats_kernel" \
    -s 1000 -c 14 -o /tmp/prof_full $CMD > gpurun_out/ncu/ncu2.log 2>&1
ncu -i /tmp/prof_full.ncu-rep --page raw --csv > gpurun_out/ncu/full_raw.csv 2> gpurun_out/ncu/export.log
ls -la /tmp/prof_full.ncu-rep gpurun_out/ncu/
