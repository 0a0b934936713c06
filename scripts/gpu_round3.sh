#!/bin/bash
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest.log 2>&1
cat gpurun_out/pytest.log
for f in 1 0; do
  echo "== MCN_BN_RUNS=$f"
  MCN_BN_RUNS=$f timeout 120 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_runs$f.json 2> gpurun_out/bench_runs$f.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
