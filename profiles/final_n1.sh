#!/bin/bash
# Round-2 single-GPU evidence run (on the GPU box): bench line, launch list, ncu --set full of the dominant kernels,
# reference arm, 1k-step curve parity with the reference's horizon schedule.  Outputs under gpurun_out/r02_*.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/r02_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cf-phase 1 > gpurun_out/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_igemm_v3_kernel|conv3x3_wgrad_v2_kernel" -s 60 -c 6 \
  -o gpurun_out/r02_ncu_full_trunk python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cf-phase 1 > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none -k regex:"clip_adam|bce_finalize|pack_nchw|plane_colsum|wgrad_reduce" -s 40 -c 10 \
  -o gpurun_out/r02_ncu_full_membound python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cf-phase 1 > gpurun_out/r02_ncu_full2.log 2>&1; echo "ncu full2 rc=$?"
python profiles/curve_parity.py --steps 1000 --horizon 10 --schedule --out gpurun_out/r02_curve_parity.json > gpurun_out/r02_curve_parity.log 2>&1; echo "curve rc=$?"
tail -14 gpurun_out/r02_curve_parity.log
