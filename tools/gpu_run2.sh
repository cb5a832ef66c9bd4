#!/usr/bin/env bash
# Multi-GPU pass: allreduce parity (in-process peers + cross-process IPC) and the bandwidth sweep.
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_allreduce_gpu.py tests/test_allreduce_ipc.py -m gpu -x -q > gpurun_out/pytest_ar_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ar_$N.log
tail -25 gpurun_out/pytest_ar_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sweep_allreduce.py --max-bytes $((1<<28)) > gpurun_out/sweep_ar_$N.log 2>&1
grep '^{' gpurun_out/sweep_ar_$N.log | tail -60
grep -v '^{' gpurun_out/sweep_ar_$N.log | tail -20
