#!/usr/bin/env bash
# round 2, run 3: HP-B tests + N=1 bench on one GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batcher_gpu.py tests/test_copy_gpu.py tests/test_env_workers_gpu.py -x -q > gpurun_out/r02_pytest_hpb.log 2>&1
tail -12 gpurun_out/r02_pytest_hpb.log
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/r02_bench_1gpu.log 2>&1
grep '^{' gpurun_out/r02_bench_1gpu.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print(json.dumps(d['roofline']['per_op']))
print(d['roofline_nvlink'])
print(d['step_ms'], d['loop_stats_rank0'])
"
grep -i -E "error|Traceback" -A5 gpurun_out/r02_bench_1gpu.log | head -30
