#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$N.log
tail -6 gpurun_out/pytest_gpu_$N.log
timeout 900 python bench.py --gpus 1 --steps 30 --warmup 8 > gpurun_out/bench_1.log 2>&1; tail -1 gpurun_out/bench_1.log | cut -c1-3500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 8 > gpurun_out/bench_$N.log 2>&1; tail -1 gpurun_out/bench_$N.log | cut -c1-3000
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sweep_allreduce.py --max-bytes $((1<<28)) > gpurun_out/sweep_ar_$N.log 2>&1
grep '^{' gpurun_out/sweep_ar_$N.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['n_gpus'], r['bytes'], r['algo'], r.get('round_us'), r['kernel_us'], r['busbw_gbs'], r.get('exact'), r.get('kernel_us_min_med_max'))"
