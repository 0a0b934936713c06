#!/bin/bash
# ncu --set full on single harness cases (small memory footprint -> fast replays)
mkdir -p gpurun_out/ncuh
for c in "$@"; do
  build/tc_harness $c > gpurun_out/ncuh/plain_$c.log 2>&1 &&
  ncu --set full --clock-control none -k regex:"halo_conv_kernel|gemm_conv_kernel|wgrad_kernel" -s 3 -c 1 \
      -o /tmp/h_$c build/tc_harness $c > gpurun_out/ncuh/ncu_$c.log 2>&1
  ncu -i /tmp/h_$c.ncu-rep --page details --csv > gpurun_out/ncuh/details_$c.csv 2>/dev/null
  ncu -i /tmp/h_$c.ncu-rep --page raw --csv > gpurun_out/ncuh/raw_$c.csv 2>/dev/null
done
ls -la gpurun_out/ncuh
