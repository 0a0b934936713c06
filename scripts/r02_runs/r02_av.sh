#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "resnet50_conv_shapes or tensor_core or wsgn or transposed" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_av0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_av0.log | cut -c1-400 | head -20
for v in new; do
timeout 600 python bench.py --no-cpu-baseline --steps 20 --profile-json gpurun_out/prof_r02av_$v.json 2> gpurun_out/bench_r02av_$v.err > gpurun_out/bench_r02av_$v.json
grep "timed region" gpurun_out/bench_r02av_$v.err | tail -1
python -c "
import json;d=json.load(open('gpurun_out/prof_r02av_$v.json'))
print({k[:14]:round(v['ms'],3) for k,v in d['classes'].items() if k.startswith('wgrad')})"
done
