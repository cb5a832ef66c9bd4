#!/usr/bin/env bash
# First GPU pass: parity tests, copy sweep for both implementations, launch list + one full ncu capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep_copy.json
for impl in ldg tma; do
  MB_COPY_IMPL=$impl timeout 600 python tools/sweep_copy.py --envs 64 256 1024 4096 --reps 10 > gpurun_out/sweep_$impl.log 2>&1
  tail -30 gpurun_out/sweep_$impl.log
done
# reference built with -DUSE_CUDA: does it import and batch on the GPU?
PYTHONPATH=oracle/_ref_cuda timeout 120 python -c "
import moolib, torch
b = moolib.Batcher(3, device='cuda:0')
for i in range(3): b.stack(torch.ones(2, device='cuda:0')*i)
print('ref cuda batcher', b.get())
" > gpurun_out/ref_cuda.log 2>&1; tail -3 gpurun_out/ref_cuda.log
for impl in ldg tma; do
MB_COPY_IMPL=$impl timeout 300 ncu --set full --clock-control none --import-source on -k regex:copy2d -s 30 -c 2 -o gpurun_out/prof_copy_$impl -f \
  python tools/sweep_copy.py --envs 1024 --reps 1 --warmup 1 --out gpurun_out/ncu_dummy.json > gpurun_out/ncu_$impl.log 2>&1
done
ls -la gpurun_out
