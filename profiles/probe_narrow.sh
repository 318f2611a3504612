#!/bin/bash
# Ablation + ncu probe of the narrow-channel conv kernels (16 -> 128 "expand", 256 -> 16 / 128 -> 16 heads).
# Writes gpurun_out/r02_probe_*.  Run on the GPU box:  bash profiles/probe_narrow.sh
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_ablate.log
: > $OUT
for d in 0 1 2 4 8 64 5 13 15; do
  echo "=== SCMGAN_DEBUG=$d" >> $OUT
  SCMGAN_DEBUG=$d python profiles/microbench_layers.py --only "tr conv1" 2>&1 | grep "us " >> $OUT
done
for d in 0 1 2 3; do
  echo "=== heads SCMGAN_DEBUG=$d" >> $OUT
  SCMGAN_DEBUG=$d python profiles/microbench_layers.py --only "tr conv6" 2>&1 | grep "us " >> $OUT
  SCMGAN_DEBUG=$d python profiles/microbench_layers.py --only "tr dz" 2>&1 | grep "us " >> $OUT
done
SCMGAN_DEBUG=4096 python profiles/expand_timeline.py --batch 32 --n 128 > gpurun_out/r02_probe_timeline.log 2>&1
python profiles/microbench_layers.py --json gpurun_out/r02_layers.json > gpurun_out/r02_layers.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_expand|igemm_v2" -c 12 -o gpurun_out/r02_probe_narrow \
  python profiles/microbench_layers.py --only "tr " > gpurun_out/r02_probe_ncu.log 2>&1
echo done
