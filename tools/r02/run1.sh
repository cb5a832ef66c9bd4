#!/usr/bin/env bash
# round 2, run 1: gated allreduce + zero-copy accumulator on N GPUs: tests, smoke, short bench
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_${N}gpu.log 2>&1
tail -15 gpurun_out/r02_pytest_${N}gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_${N}gpu.log 2>&1
tail -3 gpurun_out/r02_smoke_${N}gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_1gpu.log 2>&1
grep '^{' gpurun_out/r02_bench_1gpu.log | tail -1 | cut -c1-3000
grep -i -E "error|Traceback" -A5 gpurun_out/r02_bench_1gpu.log | head -30
if [ "$N" -gt 1 ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.log 2>&1
  grep '^{' gpurun_out/r02_bench_${N}gpu.log | tail -1 | cut -c1-3000
  grep -i -E "error|Traceback" -A5 gpurun_out/r02_bench_${N}gpu.log | head -30
fi
