#!/usr/bin/env bash
# round 2, 8-GPU pass: multi-GPU parity tests, smoke on all GPUs, bench at N=8/4/2, configs[4] GPU leg (K-A2 micro-benchmark to 1 GB)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/r02_nvidia_smi_topo_${N}gpu.txt 2>&1
timeout 420 python -m pytest tests/test_gated_gpu.py tests/test_allreduce_gpu.py tests/test_allreduce_ipc.py tests/test_accumulator_gpu.py -q > gpurun_out/r02_pytest_ar_${N}gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_ar_${N}gpu.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_${N}gpu.log 2>&1
tail -1 gpurun_out/r02_smoke_${N}gpu.log
for W in $N 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2951$W bench.py --gpus $W --steps 40 --warmup 8 > gpurun_out/r02_bench_${W}gpu.log 2>&1
  grep '^{' gpurun_out/r02_bench_${W}gpu.log | tail -1 > gpurun_out/r02_bench_${W}gpu.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${W}gpu.json'))
    print($W, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['parity']['exact'])
    print(d['roofline_nvlink']); print(d['step_ms'], d['slowest_rank'], d['frames_per_opt_step'])
except Exception as ex:
    print('bench $W failed', ex)
PY
  grep -i -E "error|Traceback" -A5 gpurun_out/r02_bench_${W}gpu.log | head -12
done
python tools/r02/ar_micro.py --tag n8 --sizes 4096,65536,1048576,4377904,16777216,67108864,268435456,1073741824 --iters 20 > gpurun_out/r02_ar_micro_${N}gpu.jsonl 2>gpurun_out/r02_ar_micro_err.log
python tools/r02/ar_micro.py --tag n4 --world 4 --sizes 4096,65536,1048576,4377904,16777216,67108864,268435456,1073741824 --iters 20 > gpurun_out/r02_ar_micro_4gpu.jsonl 2>>gpurun_out/r02_ar_micro_err.log
python tools/r02/ar_micro.py --tag n2 --world 2 --sizes 4096,65536,1048576,4377904,16777216,67108864,268435456,1073741824 --iters 20 --algos oneshot > gpurun_out/r02_ar_micro_2gpu.jsonl 2>>gpurun_out/r02_ar_micro_err.log
cat gpurun_out/r02_ar_micro_${N}gpu.jsonl gpurun_out/r02_ar_micro_4gpu.jsonl gpurun_out/r02_ar_micro_2gpu.jsonl | cut -c1-260; tail -3 gpurun_out/r02_ar_micro_err.log
