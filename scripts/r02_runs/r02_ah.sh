#!/bin/bash
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 240 $TR scripts/check_dp.py bf16 > gpurun_out/check_dp_bf16_ah.log 2>&1; echo "check bf16 rc=$?"; grep -E "PASS|FAIL|Error|error|timeout" gpurun_out/check_dp_bf16_ah.log | cut -c1-300 | head -12
for v in base cta4 cta16; do
unset NCCL_MAX_CTAS
if [ $v = cta4 ]; then export NCCL_MAX_CTAS=4; fi
if [ $v = cta16 ]; then export NCCL_MAX_CTAS=16; fi
timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ah_$v.json 2> gpurun_out/bench_ah_$v.err; echo "bench $v rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_ah_$v.json'));print('$v', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['config']['sync_bn_exchange'][:20], d['config']['grad_allreduce'], (d.get('dp_parity') or {}).get('replicas_identical'), (d.get('dp_parity') or {}).get('max_var_diff'))" || tail -5 gpurun_out/bench_ah_$v.err
done
