#!/usr/bin/env bash
# ncu evidence (1 GPU): launch list of the bench command + full captures of the dominant kernels
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --ref-cuda 0 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:copy2d_hybrid -s 30 -c 3 -o gpurun_out/prof_copy_hybrid -f \
  python tools/sweep_copy.py --envs 256 --reps 2 --warmup 1 --out gpurun_out/ncu_dummy.json > gpurun_out/ncu_hybrid.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ar_ -s 4 -c 4 -o gpurun_out/prof_allreduce_w1 -f \
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --ref-cuda 0 > gpurun_out/ncu_ar.log 2>&1
timeout 300 python -m pytest tests/test_env_workers_gpu.py -m gpu -x -q 2>&1 | tail -5
ls -la gpurun_out | tail -12
