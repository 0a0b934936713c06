#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_features.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_x.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_x.log | cut -c1-600 | head -20
MCN_LIB=$PWD/myconvnet_b200/libmcn_timing.so timeout 600 python scripts/role_timing.py 2> gpurun_out/role_timing_x.err | grep " w \|case" > gpurun_out/role_timing_x.txt
cat gpurun_out/role_timing_x.txt
for w in 1 0; do
MCN_WGRAD_WIDE=$w timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02x_w$w.json 2> gpurun_out/bench_r02x_w$w.err > gpurun_out/bench_r02x_w$w.json
grep "timed region" gpurun_out/bench_r02x_w$w.err
python -c "
import json;d=json.load(open('gpurun_out/prof_r02x_w$w.json'))
print({k[:12]:round(v['ms'],3) for k,v in d['classes'].items() if k[:3] in ('wgr','con')})"
done
