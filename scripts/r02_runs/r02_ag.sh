#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "wgrad or tc or conv" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_ag0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_ag0.log | cut -c1-600 | head -20
for v in new nofuse; do
if [ $v = nofuse ]; then export MCN_WGRAD_FUSE_REDUCE=0; fi
timeout 600 python bench.py --no-cpu-baseline --steps 20 --profile-json gpurun_out/prof_r02ag_$v.json 2> gpurun_out/bench_r02ag_$v.err > gpurun_out/bench_r02ag_$v.json
grep "timed region" gpurun_out/bench_r02ag_$v.err | tail -1; tail -2 gpurun_out/bench_r02ag_$v.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/prof_r02ag_$v.json'))
print({k[:12]:round(v['ms'],3) for k,v in d['classes'].items() if v['ms']>0.3})"
done
unset MCN_WGRAD_FUSE_REDUCE
timeout 900 python -m pytest tests/test_gpu_resnet.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_ag1.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_ag1.log | cut -c1-600 | head -20
