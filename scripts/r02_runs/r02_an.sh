#!/bin/bash
mkdir -p gpurun_out
for c in effnet_b0; do
timeout 600 python bench.py --config $c --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02an_$c.json 2> gpurun_out/bench_r02an_$c.err > gpurun_out/bench_r02an_$c.json
grep "timed region" gpurun_out/bench_r02an_$c.err | tail -1; tail -2 gpurun_out/bench_r02an_$c.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/prof_r02an_$c.json'))
print({k[:14]:round(v['ms'],3) for k,v in d['classes'].items() if v['ms']>0.5})"
done
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k efficientnet 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_an1.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_an1.log | cut -c1-600 | head -20
