#!/usr/bin/env bash
# round 2, run D (1 GPU): the driver's tier -- full GPU test suite, smoke, default bench (with cpu_baseline), reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_1gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_1gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu_full.log 2>&1
grep '^{' gpurun_out/r02_bench_1gpu_full.log | tail -1 > gpurun_out/r02_bench_1gpu_full.json
cp gpurun_out/r02_bench_1gpu_full.json gpurun_out/r02_bench_1gpu.json
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_1gpu_full.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}, d['e2e'], d['clocks'])
print(d['roofline']['per_op'], d['roofline']['traffic']); print(d['roofline_nvlink']); print(d['step_ms'])
print(d.get('cpu_baseline')); print(d.get('reference_cuda_model'))
PY
grep -i -E "error|Traceback" -A5 gpurun_out/r02_bench_1gpu_full.log | head
timeout 400 python bench.py --impl reference --steps 8 --warmup 3 2>/dev/null | grep '^{' | cut -c1-600
