#!/bin/bash
mkdir -p gpurun_out
for v in 100 75 50; do
export MCN_WGRAD_WAVE_PCT=$v
timeout 600 python bench.py --no-cpu-baseline --steps 20 --profile-json gpurun_out/prof_r02at_$v.json 2> gpurun_out/bench_r02at_$v.err > gpurun_out/bench_r02at_$v.json
grep "timed region" gpurun_out/bench_r02at_$v.err | tail -1
python -c "
import json;d=json.load(open('gpurun_out/prof_r02at_$v.json'))
print({k[:14]:round(v['ms'],3) for k,v in d['classes'].items() if k.startswith('wgrad')})"
done
