#!/bin/bash
# 2-GPU validation: correctness of the peer all-reduce + overlapped buckets, then bench A/B.
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 240 $TR scripts/check_dp.py f32 > gpurun_out/check_dp_f32.log 2>&1; echo "check f32 rc=$?"; grep -E "PASS|FAIL|Error|error|timeout" gpurun_out/check_dp_f32.log | cut -c1-400 | head -20
timeout 240 $TR scripts/check_dp.py bf16 > gpurun_out/check_dp_bf16.log 2>&1; echo "check bf16 rc=$?"; grep -E "PASS|FAIL|Error|error|timeout" gpurun_out/check_dp_bf16.log | cut -c1-400 | head -20
timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_dp${N}_new.json 2> gpurun_out/bench_dp${N}_new.err; echo "bench new rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_dp${N}_new.json'));print('NEW', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['sync_bn_exchange'], d['config']['grad_allreduce'])" || tail -5 gpurun_out/bench_dp${N}_new.err
MCN_PEER_ALLREDUCE=0 MCN_OVERLAP_GRADS=0 timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_dp${N}_old.json 2> gpurun_out/bench_dp${N}_old.err; echo "bench old rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_dp${N}_old.json'));print('OLD', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['sync_bn_exchange'], d['config']['grad_allreduce'])" || tail -5 gpurun_out/bench_dp${N}_old.err
