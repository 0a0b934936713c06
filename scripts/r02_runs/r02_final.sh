#!/bin/bash
# final evidence of the round: scripts/r02_profile.sh (five bench lines, role counters, ncu) + the whole GPU suite
bash scripts/r02_profile.sh
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/r02/pytest_gpu.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/r02/pytest_gpu.log | cut -c1-300 | head -20
