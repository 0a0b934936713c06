#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "stem or maxpool or pool" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_ab0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_ab0.log | cut -c1-600 | head -20
ONLY_STEM=1 MCN_LIB=$PWD/myconvnet_b200/libmcn_timing.so timeout 600 python scripts/role_timing.py 2> gpurun_out/role_timing_ab.err | tee gpurun_out/role_timing_ab.txt
tail -3 gpurun_out/role_timing_ab.err
for v in new; do
timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02ab_$v.json 2> gpurun_out/bench_r02ab_$v.err > gpurun_out/bench_r02ab_$v.json
grep "timed region" gpurun_out/bench_r02ab_$v.err | tail -1; tail -2 gpurun_out/bench_r02ab_$v.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/prof_r02ab_$v.json'))
print({k[:12]:round(v['ms'],3) for k,v in d['classes'].items() if v['ms']>0.3})
print([(l['tag'],round(l['us'],1)) for l in d['launches'] if 'max_pool' in l['tag'] or 'block_0/conv_0' in l['tag']])"
done
