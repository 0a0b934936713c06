#!/bin/bash
# Round profile of one eager training step (the NVTX range "mcn_profiled_step" in bench.py):
#  (1) launch list with duration and DRAM bytes of EVERY launch of the step,
#  (2) --set full on the first launches of each heavy kernel, exported to CSV on the box.
# Each ncu pass only after the identical plain command exited 0 (B200_PROFILING.md).
mkdir -p gpurun_out/ncu
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --stall-timeout 100000"
$CMD > gpurun_out/ncu/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu/plain.log; exit 1; }
ncu --nvtx --nvtx-include "mcn_profiled_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 400 --csv --log-file gpurun_out/ncu/launches.csv $CMD > gpurun_out/ncu/ncu1.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/ncu/launches.csv)"
ncu --nvtx --nvtx-include "mcn_profiled_step/" --set full --clock-control none \
    -k regex:"gemm_conv_kernel|halo_conv_kernel|wgrad_halo_kernel|wgrad_kernel|stem_fprop_kernel|stem_wgrad_kernel|bn_bwd_apply_pipe_kernel|bn_bwd_reduce_kernel|bn_apply_pipe_kernel|maxpool_fwd_tap|maxpool_bwd_tap" \
    -s 6 -c 40 -o /tmp/prof_full $CMD > gpurun_out/ncu/ncu2.log 2>&1
echo "full set rc=$?"
ncu -i /tmp/prof_full.ncu-rep --page raw --csv > gpurun_out/ncu/full_raw.csv 2> gpurun_out/ncu/export.log
# the same for the start of the backward pass (block_4 / block_3: dgrad with fused BN sums, wgrad, masked BN passes)
ncu --nvtx --nvtx-include "mcn_profiled_step/" --set full --clock-control none \
    -k regex:"gemm_conv_kernel|halo_conv_kernel|wgrad_halo_kernel|wgrad_kernel|bn_bwd_apply_pipe_kernel|bn_bwd_reduce_kernel|maxpool_bwd_tap" \
    -s 112 -c 48 -o /tmp/prof_full_bwd $CMD > gpurun_out/ncu/ncu3.log 2>&1
echo "full set (backward) rc=$?"
ncu -i /tmp/prof_full_bwd.ncu-rep --page raw --csv > gpurun_out/ncu/full_raw_bwd.csv 2>> gpurun_out/ncu/export.log
ls -la /tmp/prof_full.ncu-rep gpurun_out/ncu/ | tail -8
