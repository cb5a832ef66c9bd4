#!/usr/bin/env bash
# round 2, run C (1 GPU): new tests, ncu of the unroll gather (large-parameter kernel) and of the learner ops, launch list of the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_1gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_1gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:copy2d_hybrid_kernel_l -s 1 -c 2 -o gpurun_out/r02_prof_gather -f python tools/r02/gather_micro.py --reps 4 > gpurun_out/r02_ncu_gather.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 9000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 6 --warmup 10 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/r02_bench_1gpu.log 2>&1
grep '^{' gpurun_out/r02_bench_1gpu.log | tail -1 > gpurun_out/r02_bench_1gpu.json
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print(d['roofline']['per_op']); print(d['step_ms'])
PY
grep -i -E "error|Traceback" -A5 gpurun_out/r02_bench_1gpu.log | head
