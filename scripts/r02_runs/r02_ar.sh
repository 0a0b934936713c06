#!/bin/bash
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() {
v=$1
MCN_PEER_TIMEOUT_S=30 timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ar_$v.json 2> gpurun_out/bench_ar_$v.err; echo "bench $v rc=$?"
python - <<PY
import json
for line in open('gpurun_out/bench_ar_$v.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print('$v', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['config']['grad_allreduce'], (d.get('dp_parity') or {}).get('pass'))
PY
}
MCN_NCCL_NVLS=1 run nvls
MCN_OVERLAP_GRADS=0 MCN_BUCKET_ELEMS=33554432 run one_bucket_end
