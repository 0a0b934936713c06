#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_sd.py 2>&1 | grep -v "shape\|drop rate" > gpurun_out/debug_sd.txt
cat gpurun_out/debug_sd.txt | cut -c1-300
(time timeout 1500 python -m pytest tests/test_gpu_features.py tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/pytest_f.log 2>&1
cat gpurun_out/pytest_f.log | cut -c1-800
timeout 600 python bench.py --config effnet_b0 --no-cpu-baseline --steps 8 --profile-json gpurun_out/prof_r02f_effnet_b0.json 2> gpurun_out/bench_r02f_effnet.err > gpurun_out/bench_r02f_effnet.json
cat gpurun_out/bench_r02f_effnet.json | cut -c1-200
