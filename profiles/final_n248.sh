#!/bin/bash
# bench.py at N GPUs of one box, with the gradient exchange and - as a diagnostic - without it (SCMGAN_DP_NOSYNC=1: every
# rank then runs at its own pace and bench.py's rank_ms_per_step shows the spread the exchange has to wait for).
#   gpurun --gpus 8 -- bash profiles/final_n248.sh 8
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench n$N rc=$?"
SCMGAN_DP_NOSYNC=1 $TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}_nosync.json 2> gpurun_out/r02_bench_n${N}_nosync.err; echo "nosync rc=$?"
python -c "
import json,sys
for f in ('gpurun_out/r02_bench_n$N.json','gpurun_out/r02_bench_n${N}_nosync.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d.get('rank_ms_per_step'))
    except Exception as e: print(f, 'unreadable', e)
"
