#!/usr/bin/env bash
# round 2, run A (1 GPU): full GPU test suite, configs[3] sweep with reference CPU legs, ncu captures, bench N=1,
# reference allreduce CPU leg (host cores of the GPU box)
mkdir -p gpurun_out
nproc > gpurun_out/r02_host_cores.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_1gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_1gpu.log
timeout 600 python tools/r02/sweep_hpb.py > gpurun_out/r02_sweep_hpb.jsonl 2> gpurun_out/r02_sweep_hpb.err
cat gpurun_out/r02_sweep_hpb.jsonl | cut -c1-900; tail -2 gpurun_out/r02_sweep_hpb.err
# ncu: launch list of a short bench, then --set full of the gather launch and of the N=1 reduce kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:copy2d_hybrid -s 4 -c 2 -o gpurun_out/r02_prof_gather -f python tools/r02/gather_micro.py --reps 4 > gpurun_out/r02_ncu_gather.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ar_ -c 6 -o gpurun_out/r02_prof_ar1 -f python tools/r02/ar_micro.py --world 1 --sizes 4377904 --algos oneshot --iters 2 > gpurun_out/r02_ncu_ar1.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 40 --warmup 8 > gpurun_out/r02_bench_1gpu_full.log 2>&1
grep '^{' gpurun_out/r02_bench_1gpu_full.log | tail -1 > gpurun_out/r02_bench_1gpu_full.json
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_1gpu_full.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d.get('cpu_baseline'), d.get('reference_cuda_model'))
PY
timeout 400 python tools/r02/sweep_ref_allreduce.py --budget-s 6 --sizes 1024 65536 1048576 4377904 16777216 67108864 268435456 > gpurun_out/r02_sweep_ref_allreduce.jsonl 2> gpurun_out/r02_sweep_ref_allreduce.err
cut -c100-400 gpurun_out/r02_sweep_ref_allreduce.jsonl; tail -2 gpurun_out/r02_sweep_ref_allreduce.err
