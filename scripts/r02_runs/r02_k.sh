#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_k.log
grep -n "Error\|assert\|FAILED\|passed\|failed" gpurun_out/pytest_k.log | cut -c1-1200 | head -60
