#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "fused_bn_backward" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_y0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_y0.log | cut -c1-600 | head -20
for f in 1 0; do
MCN_FUSE_BN_BWD=$f timeout 600 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r02y_f$f.json 2> gpurun_out/bench_r02y_f$f.err > gpurun_out/bench_r02y_f$f.json
grep "timed region" gpurun_out/bench_r02y_f$f.err; tail -3 gpurun_out/bench_r02y_f$f.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/prof_r02y_f$f.json'))
print({k[:12]:round(v['ms'],3) for k,v in d['classes'].items()})"
done
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_y.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_y.log | cut -c1-600 | head -30
