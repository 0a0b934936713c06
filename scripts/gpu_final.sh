#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6) | cut -c1-400
timeout 150 python bench.py --profile-json gpurun_out/prof_final.json > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
grep "^{" gpurun_out/bench_final.json | cut -c1-2500
timeout 120 build/stream_harness 4 > gpurun_out/stream2.log 2>&1; tail -16 gpurun_out/stream2.log
