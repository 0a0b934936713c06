#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_resnet.py "tests/test_gpu_features.py::test_stochastic_depth_and_dropout_in_a_step" -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/pytest_h.log
grep -n "Error\|assert\|FAILED\|passed\|failed" gpurun_out/pytest_h.log | cut -c1-900 | head -80
