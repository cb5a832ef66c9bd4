#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
MOOLIB_B200_TRACE=1 BENCH_DEBUG=600 timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_dbg_$N.log 2>&1
grep -v "^\[W\|^W0" gpurun_out/bench_dbg_$N.log | tail -80 | cut -c1-300
