#!/bin/bash
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
MCN_PEER_TIMEOUT_S=20 timeout 240 $TR scripts/check_dp.py bf16 > gpurun_out/check_dp_bf16_aj.log 2>&1; echo "check bf16 rc=$?"; grep -E "PASS|FAIL|Error|error|timeout" gpurun_out/check_dp_bf16_aj.log | cut -c1-200 | head -12
run() {
v=$1
MCN_PEER_TIMEOUT_S=20 timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_aj_$v.json 2> gpurun_out/bench_aj_$v.err; echo "bench $v rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_aj_$v.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('$v', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['config']['grad_allreduce'], (d.get('dp_parity') or {}).get('pass'))
PY
}
run ll
MCN_PEER_LL=0 run flag
MCN_BUCKET_ELEMS=8388608 run ll_8m
