#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "loss or resize or xent or softmax" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_ak0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_ak0.log | cut -c1-600 | head -20
for v in new old; do
if [ $v = old ]; then export MCN_XENT_ROWS=0 MCN_RESIZE_TABLE=0; fi
timeout 600 python bench.py --config deeplab_r50_512 --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02ak_$v.json 2> gpurun_out/bench_r02ak_$v.err > gpurun_out/bench_r02ak_$v.json
grep "timed region" gpurun_out/bench_r02ak_$v.err | tail -1; tail -2 gpurun_out/bench_r02ak_$v.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/prof_r02ak_$v.json'))
print({k[:14]:round(v['ms'],3) for k,v in d['classes'].items() if v['ms']>0.3})"
done
unset MCN_XENT_ROWS MCN_RESIZE_TABLE
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "deeplab or segnet" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_ak1.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_ak1.log | cut -c1-600 | head -20
