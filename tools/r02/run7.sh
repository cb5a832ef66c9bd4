#!/usr/bin/env bash
# pacing experiment: bench at N = all GPUs of the box and N = 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for W in $N 1; do
  if [ "$W" -gt 1 ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2952$W bench.py --gpus $W --steps 40 --warmup 8 > gpurun_out/r02_bench_${W}gpu.log 2>&1
  else
    timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/r02_bench_${W}gpu.log 2>&1
  fi
  grep '^{' gpurun_out/r02_bench_${W}gpu.log | tail -1 > gpurun_out/r02_bench_${W}gpu.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${W}gpu.json'))
    print($W, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['parity']['exact'])
    print(d['roofline']['per_op']); print(d['roofline_nvlink']); print(d['step_ms'], d['slowest_rank'], d['frames_per_opt_step'])
    print(d['loop_stats_rank0'])
except Exception as ex:
    print('bench $W failed', ex)
PY
  grep -i -E "error|Traceback" -A5 gpurun_out/r02_bench_${W}gpu.log | head -20
done
