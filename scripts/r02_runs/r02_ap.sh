#!/bin/bash
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
MCN_PEER_TIMEOUT_S=30 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ap_8.json 2> gpurun_out/bench_ap_8.err; echo "bench rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_ap_8.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('N=8', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['config']['grad_allreduce'], (d.get('dp_parity') or {}).get('pass'), d['config']['sync_bn_exchange'][:30])
PY
tail -4 gpurun_out/bench_ap_8.err | cut -c1-200
