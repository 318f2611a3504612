#!/bin/bash
# How much of the data-parallel step is the exchange itself?  bench.py at N GPUs with the exchange, without it (rank-skew
# diagnostic), and with NCCL limited to a few CTAs (its kernels share the SMs with the encoder's backward).
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # label, env...
  label=$1; shift
  env "$@" $TR --master-port 295$((RANDOM % 90 + 10)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_n${N}_$label.json 2> gpurun_out/r02_n${N}_$label.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r02_n${N}_$label.json').read().strip().splitlines()[-1]); print('N=$N $label', round(d['ms_per_step'],4), d.get('rank_ms_per_step'))
except Exception as e: print('N=$N $label unreadable', e)
"
}
run default X=1
run nosync SCMGAN_DP_NOSYNC=1
run maxctas2 NCCL_MAX_CTAS=2
run maxctas4 NCCL_MAX_CTAS=4
run maxctas8 NCCL_MAX_CTAS=8
