#!/bin/bash
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest.log 2>&1
cat gpurun_out/pytest.log | cut -c1-600
timeout 120 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_r4.json 2> gpurun_out/bench_r4.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'])"
