#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_resnet.py -m gpu -q --no-header -p no:cacheprovider --tb=short -x -k "pool or argmax or config_1 or baseline" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_ax0.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_ax0.log | cut -c1-400 | head -20
timeout 600 python bench.py --no-cpu-baseline --steps 20 --profile-json gpurun_out/prof_r02ax.json 2> gpurun_out/bench_r02ax.err > gpurun_out/bench_r02ax.json
grep "timed region" gpurun_out/bench_r02ax.err | tail -1
python -c "
import json;d=json.load(open('gpurun_out/prof_r02ax.json'))
print([(l['tag'],round(l['us'],1),round(l['GB/s'])) for l in d['launches'] if 'max_pool' in l['tag']])"
