#!/bin/bash
mkdir -p gpurun_out/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --stall-timeout 100000"
ncu --nvtx --nvtx-include "mcn_profiled_step/" --set full --import-source on --clock-control none \
    -k regex:"maxpool_fwd_tap|maxpool_bwd_tap" -c 2 -o /tmp/prof_pool $CMD > gpurun_out/ncu/ncu_pool.log 2>&1
echo "rc=$?"
ncu -i /tmp/prof_pool.ncu-rep --page raw --csv > gpurun_out/ncu/pool_raw.csv 2>> gpurun_out/ncu/export.log
ncu -i /tmp/prof_pool.ncu-rep --page details > gpurun_out/ncu/pool_details.txt 2>> gpurun_out/ncu/export.log
wc -c gpurun_out/ncu/pool_raw.csv gpurun_out/ncu/pool_details.txt
