#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -q --no-header -p no:cacheprovider --tb=short -k "group_norm or standardisation or wsgn" 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_v.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/pytest_v.log | cut -c1-900 | head -40
