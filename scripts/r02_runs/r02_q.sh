#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_q1.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_q1.log | cut -c1-500 | head -20
MCN_LIB=$PWD/myconvnet_b200/libmcn_timing.so timeout 600 python scripts/role_timing.py 2> gpurun_out/role_timing_q.err | grep " w " > gpurun_out/role_timing_q.txt
cat gpurun_out/role_timing_q.txt; tail -5 gpurun_out/role_timing_q.err
for r in k all; do
MCN_FUSE_STATS_RULE=$r timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02q_$r.json 2> gpurun_out/bench_r02q_$r.err > gpurun_out/bench_r02q_$r.json
cut -c1-200 gpurun_out/bench_r02q_$r.json
done
# ncu: a handful of small BN launches and conv launches, full sets
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "mcn_profiled_step/" \
  --kernel-name regex:"bn_bwd_reduce_kernel|bn_bwd_apply_kernel|bn_apply_kernel" --launch-skip 100 --launch-count 12 \
  -o gpurun_out/r02q_bn -f python bench.py --no-cpu-baseline --steps 1 --warmup 1 --no-graph > gpurun_out/ncu_q.log 2>&1
tail -3 gpurun_out/ncu_q.log
