#!/bin/bash
mkdir -p gpurun_out
for v in base all; do
if [ $v = all ]; then export MCN_FUSE_STATS_RULE=all; fi
timeout 600 python bench.py --no-cpu-baseline --steps 20 --profile-json gpurun_out/prof_r02ao_$v.json 2> gpurun_out/bench_r02ao_$v.err > gpurun_out/bench_r02ao_$v.json
grep "timed region" gpurun_out/bench_r02ao_$v.err | tail -1; tail -1 gpurun_out/bench_r02ao_$v.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/prof_r02ao_$v.json'))
print({k[:14]:round(v['ms'],3) for k,v in d['classes'].items() if v['ms']>0.3})"
done
