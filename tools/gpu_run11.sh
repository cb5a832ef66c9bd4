#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_allreduce_ipc.py "tests/test_allreduce_gpu.py::test_many_rounds_two_slots_exact_integers" "tests/test_allreduce_gpu.py::test_atari_grad_list_matches_oracle" -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sweep_allreduce.py --sizes 4096 65536 1048576 4377904 16777216 67108864 268435456 > gpurun_out/sweep_ar_$N.log 2>&1
grep '^{' gpurun_out/sweep_ar_$N.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['n_gpus'], r['bytes'], r['algo'], r.get('round_us'), r['kernel_us'], r['busbw_gbs'], r.get('exact'), r.get('kernel_us_min_med_max'))"
grep -v '^{' gpurun_out/sweep_ar_$N.log | grep -i -E "error|Traceback" | head -5
