#!/usr/bin/env bash
# final 1-GPU check: what the driver runs at round end + the ncu capture of the big Batcher.cat launch
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -4 gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_1.log 2>&1; tail -1 gpurun_out/bench_1.log | cut -c1-4000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:copy2d_hybrid -s 58 -c 2 -o gpurun_out/prof_copy_cat -f \
  python tools/sweep_copy.py --envs 256 --reps 1 --warmup 1 --out gpurun_out/ncu_dummy.json > gpurun_out/ncu_cat.log 2>&1
ls -la gpurun_out/prof_copy_cat.ncu-rep
