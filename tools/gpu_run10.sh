#!/usr/bin/env bash
# N-GPU validation: multi-GPU parity tests, bench at N, allreduce sweep at N
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_allreduce_gpu.py tests/test_allreduce_ipc.py tests/test_accumulator_gpu.py -m gpu -x -q > gpurun_out/pytest_ar_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ar_$N.log
tail -6 gpurun_out/pytest_ar_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/bench_$N.log 2>&1; tail -1 gpurun_out/bench_$N.log | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print({k: r[k] for k in ('value','ms_per_step','e2e','gpu_launches','frames_per_opt_step','loop_stats_rank0')})"
grep -i -E "error|Traceback" gpurun_out/bench_$N.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sweep_allreduce.py --max-bytes $((1<<28)) > gpurun_out/sweep_ar_$N.log 2>&1
grep '^{' gpurun_out/sweep_ar_$N.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['n_gpus'], r['bytes'], r['algo'], r.get('round_us'), r['kernel_us'], r['busbw_gbs'], r.get('exact'), r.get('kernel_us_min_med_max'))"
grep -v '^{' gpurun_out/sweep_ar_$N.log | grep -i -E "error|Traceback" | head -5
