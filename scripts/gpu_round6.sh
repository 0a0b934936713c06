#!/bin/bash
mkdir -p gpurun_out/ncuh
MCN_WEIGHT_STATIONARY=2 timeout 200 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "deeplab and bf16" 2>&1 | grep -E "Error|error|libmcn" | head -5 | cut -c1-600
for c in 61 38; do
  ncu --set full --clock-control none -k regex:"gemm_conv_kernel" -s 3 -c 1 -o /tmp/h_$c build/tc_harness $c > gpurun_out/ncuh/ncu_$c.log 2>&1
  ncu -i /tmp/h_$c.ncu-rep --page details --csv > gpurun_out/ncuh/details_$c.csv 2>/dev/null
  ncu -i /tmp/h_$c.ncu-rep --page raw --csv > gpurun_out/ncuh/raw_$c.csv 2>/dev/null
done
ls -la gpurun_out/ncuh | tail -5
