#!/bin/bash
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
MCN_PEER_TIMEOUT_S=20 timeout 240 $TR scripts/check_dp.py bf16 > gpurun_out/check_dp_bf16_aq.log 2>&1; echo "check bf16 rc=$?"; grep -E "PASS|FAIL|Error|error|timeout" gpurun_out/check_dp_bf16_aq.log | cut -c1-160 | head -14
MCN_PEER_TIMEOUT_S=20 timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_aq.json 2> gpurun_out/bench_aq.err; echo "bench rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_aq.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('N=2', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['dp_parity'])
PY
