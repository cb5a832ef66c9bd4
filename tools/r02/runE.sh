#!/usr/bin/env bash
# round 2, final check on 2 GPUs: full GPU suite, bench N=1 (driver defaults) and N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_2gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_2gpu.log | cut -c1-200
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_1gpu.log 2>&1
grep '^{' gpurun_out/r02_bench_1gpu.log | tail -1 > gpurun_out/r02_bench_1gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu.log 2>&1
grep '^{' gpurun_out/r02_bench_2gpu.log | tail -1 > gpurun_out/r02_bench_2gpu.json
python - <<PY
import json
for n in (1,2):
    d=json.load(open(f'gpurun_out/r02_bench_{n}gpu.json'))
    print(n, {k:d[k] for k in ('value','ms_per_step','steps','warmup','settle_steps')}, d['e2e']['value'], d['parity']['exact'], d['step_ms'])
PY
grep -i -E "error|Traceback" -A5 gpurun_out/r02_bench_1gpu.log gpurun_out/r02_bench_2gpu.log | head
