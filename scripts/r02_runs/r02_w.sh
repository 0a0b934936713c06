#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_features.py -m gpu -q --no-header -p no:cacheprovider --tb=short -k "pool or gap or wsgn or resize" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_w.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_w.log | cut -c1-600 | head -20
for b in 8 16 100000; do
MCN_POOL_BLOCKS_PER_SM=$b timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02w_b$b.json 2> gpurun_out/bench_r02w_b$b.err > gpurun_out/bench_r02w_b$b.json
grep "timed region" gpurun_out/bench_r02w_b$b.err
python -c "
import json;d=json.load(open('gpurun_out/prof_r02w_b$b.json'))
print({k[:12]:round(v['ms'],3) for k,v in d['classes'].items() if k[:3] in ('max','gap','ste')})"
done
