#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/debug_determinism.py f32 224 32 1000 3 2>&1 | grep -v "shape \[\|drop rate\|Warning" | tail -60 > gpurun_out/debug_det.txt
cat gpurun_out/debug_det.txt
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_n.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_n.log | cut -c1-700 | head -60
