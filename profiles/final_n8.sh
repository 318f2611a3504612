#!/bin/bash
# Round-2 8-GPU evidence run: bench.py at N = 8 (with and without the gradient exchange: rank-skew diagnostic) and the
# configuration sweep.  Outputs under gpurun_out/r02_*.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc=$?"
cut -c1-200 gpurun_out/r02_bench_n8.json
SCMGAN_DP_NOSYNC=1 $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8_nosync.json 2> gpurun_out/r02_bench_n8_nosync.err; echo "nosync rc=$?"
python -c "
import json
for f in ('gpurun_out/r02_bench_n8.json','gpurun_out/r02_bench_n8_nosync.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d.get('rank_ms_per_step'))
"
$TR --master-port 29513 profiles/sweep.py --grid pong64,minipacman_weak,minipacman_strong,sc2,tsweep_short > gpurun_out/r02_sweep_n8.jsonl 2> gpurun_out/sweep_n8.err; echo "sweep rc=$?"
python profiles/print_sweep.py gpurun_out/r02_sweep_n8.jsonl
