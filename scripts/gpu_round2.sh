#!/bin/bash
mkdir -p gpurun_out
timeout 120 build/stream_harness 4 > gpurun_out/stream.log 2>&1; cat gpurun_out/stream.log
# f32 loss-curve sensitivity: fused finalize vs separate finalize
for f in 1 0; do
  echo "== MCN_BN_FOLD_FINALIZE=$f"
  MCN_BN_FOLD_FINALIZE=$f timeout 200 python -m pytest tests/test_gpu_resnet.py -m gpu -q -k "fused_loss_curve" 2>&1 | grep -E "AssertionError: |passed|failed" | cut -c1-400
done
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "fused_bn or from_sums" 2>&1 | tail -3
for f in 1 0; do
  echo "== MCN_FUSE_BN_STATS=$f"
  MCN_FUSE_BN_STATS=$f timeout 120 python bench.py --no-cpu-baseline --steps 10 --profile-json gpurun_out/prof_fuse$f.json 2> gpurun_out/bench_fuse$f.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
