#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "fused_bn_backward" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_z0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_z0.log | cut -c1-600 | head -20
for h in default small; do
if [ $h = small ]; then export MCN_HALO_MIN_EFF=0.7 MCN_HALO_MIN_HW=100; fi
MCN_LIB=$PWD/myconvnet_b200/libmcn_timing.so timeout 600 python scripts/role_timing.py 2> gpurun_out/role_timing_z_$h.err | grep " f \| d \|case" > gpurun_out/role_timing_z_$h.txt
cat gpurun_out/role_timing_z_$h.txt | cut -c1-130
timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02z_$h.json 2> gpurun_out/bench_r02z_$h.err > gpurun_out/bench_r02z_$h.json
grep "timed region" gpurun_out/bench_r02z_$h.err | tail -1; tail -2 gpurun_out/bench_r02z_$h.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/prof_r02z_$h.json'))
print({k[:12]:round(v['ms'],3) for k,v in d['classes'].items() if v['ms']>0.3})"
done
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_resnet.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_z.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_z.log | cut -c1-600 | head -30
