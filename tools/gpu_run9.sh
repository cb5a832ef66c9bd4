#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_accumulator_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 40 --warmup 8 --no-cpu-baseline --ref-cuda 0 > gpurun_out/bench_1.log 2>&1; tail -1 gpurun_out/bench_1.log | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print({k: r[k] for k in ('value','ms_per_step','e2e','gpu_launches','frames_per_opt_step','optimizer_steps_per_s','loop_stats_rank0')}); print(r['roofline']['per_op']); print(r.get('reference_cuda_model'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/bench_$N.log 2>&1; tail -1 gpurun_out/bench_$N.log | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print({k: r[k] for k in ('value','ms_per_step','e2e','gpu_launches','frames_per_opt_step','optimizer_steps_per_s','loop_stats_rank0')})"
