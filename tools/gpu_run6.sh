#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
MOOLIB_B200_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 8 > gpurun_out/bench_dbg_$N.log 2>&1
grep -v "^\[W\|^W0" gpurun_out/bench_dbg_$N.log | tail -80 | cut -c1-300
tail -3 gpurun_out/bench_dbg_$N.log | cut -c1-2500
for cfg in "MB_AR_BLOCKS_PER_SM=2" "MB_AR_BLOCKS_PER_SM=1" "MB_AR_FORCE_U1=1" "MB_AR_BLOCKS_PER_SM=1 MB_AR_FORCE_U1=1"; do
echo "== $cfg"
env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sweep_allreduce.py --nccl 0 --sizes 4194304 8388608 16777216 33554432 67108864 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['bytes'], r['algo'], r.get('round_us'), r['kernel_us'], r['busbw_gbs'], r.get('kernel_us_min_med_max'))"
done
