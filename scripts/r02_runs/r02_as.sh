#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py tests/test_gpu_resnet.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "fused_bn_backward or resnet or step or config" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_as0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_as0.log | cut -c1-400 | head -20
for v in new launch; do
if [ $v = launch ]; then export MCN_BN_BWD_FINALIZE_LAUNCH=1; fi
timeout 600 python bench.py --no-cpu-baseline --steps 20 --profile-json gpurun_out/prof_r02as_$v.json 2> gpurun_out/bench_r02as_$v.err > gpurun_out/bench_r02as_$v.json
grep "timed region" gpurun_out/bench_r02as_$v.err | tail -1; tail -1 gpurun_out/bench_r02as_$v.err | cut -c1-300
done
unset MCN_BN_BWD_FINALIZE_LAUNCH
MCN_LIB=$PWD/myconvnet_b200/libmcn_timing.so timeout 600 python scripts/role_timing.py > gpurun_out/role_timing_as.txt 2> gpurun_out/role_timing_as.err
tail -4 gpurun_out/role_timing_as.txt | cut -c1-250
