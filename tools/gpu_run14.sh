#!/usr/bin/env bash
# information only: the unmodified reference (USE_CUDA build) as N processes, one per GPU, through the same loop
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
MB_REF_CUDA=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --impl reference --gpus $N --steps 40 --warmup 8 --max-seconds 80 > gpurun_out/bench_refcuda_$N.log 2>&1
grep '^{' gpurun_out/bench_refcuda_$N.log | tail -1 | cut -c1-1500
grep -i -E "error|Traceback" gpurun_out/bench_refcuda_$N.log | head -5
