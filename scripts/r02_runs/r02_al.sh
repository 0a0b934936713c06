#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_features.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "tensor_core or stats or resize or conv" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_al0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_al0.log | cut -c1-600 | head -20
for v in new old; do
if [ $v = old ]; then export MCN_TMA_STORE_MIN64=1; fi
timeout 600 python bench.py --config effnet_b0 --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02al_$v.json 2> gpurun_out/bench_r02al_$v.err > gpurun_out/bench_r02al_$v.json
grep "timed region" gpurun_out/bench_r02al_$v.err | tail -1; tail -2 gpurun_out/bench_r02al_$v.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/prof_r02al_$v.json'))
print({k[:14]:round(v['ms'],3) for k,v in d['classes'].items() if v['ms']>0.5})"
done
unset MCN_TMA_STORE_MIN64
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "efficientnet" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_al1.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_al1.log | cut -c1-600 | head -20
