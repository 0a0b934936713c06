#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "stem or maxpool or pool" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_aa0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_aa0.log | cut -c1-600 | head -20
for v in new oldpool; do
if [ $v = oldpool ]; then export MCN_POOL_STRIP=0; fi
timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02aa_$v.json 2> gpurun_out/bench_r02aa_$v.err > gpurun_out/bench_r02aa_$v.json
grep "timed region" gpurun_out/bench_r02aa_$v.err | tail -1; tail -2 gpurun_out/bench_r02aa_$v.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/prof_r02aa_$v.json'))
print({k[:12]:round(v['ms'],3) for k,v in d['classes'].items() if v['ms']>0.3})
print([(l['tag'],round(l['us'],1)) for l in d['launches'] if 'max_pool' in l['tag'] or 'block_0/conv_0' in l['tag']])"
done
