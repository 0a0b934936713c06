#!/bin/bash
# last GPU call of the round: smoke(), the whole GPU suite, then the five bench lines again
mkdir -p gpurun_out/r02b
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|UserWarning" > gpurun_out/r02b/pytest_gpu.log
grep -n "Error\|assert \|^E  \|FAILED\|passed\|failed" gpurun_out/r02b/pytest_gpu.log | cut -c1-300 | head -20
timeout 900 python bench.py --steps 20 --warmup 5 --profile-json gpurun_out/r02b/kernel_classes_r50.json > gpurun_out/r02b/bench_r50.json 2> gpurun_out/r02b/bench_r50.err
cut -c1-200 gpurun_out/r02b/bench_r50.json
for c in effnet_b0 deeplab_r50_512 dcgan_64; do
timeout 600 python bench.py --config $c --no-cpu-baseline --steps 10 --profile-json gpurun_out/r02b/kernel_classes_$c.json > gpurun_out/r02b/bench_$c.json 2> gpurun_out/r02b/bench_$c.err
cut -c1-160 gpurun_out/r02b/bench_$c.json
done
