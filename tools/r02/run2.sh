#!/usr/bin/env bash
# K-A2 micro-benchmark matrix on N GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02_ar_micro_${N}gpu.jsonl
: > $OUT
python tools/r02/ar_micro.py --tag default >> $OUT 2>gpurun_out/r02_ar_micro_err.log
for U in 1 2 4 8; do
  MB_AR_UNROLL=$U python tools/r02/ar_micro.py --tag U$U --sizes 65536,4377904,67108864 >> $OUT 2>>gpurun_out/r02_ar_micro_err.log
done
MB_AR_BLOCKS_PER_SM=2 python tools/r02/ar_micro.py --tag bps2 --sizes 65536,4377904,67108864 >> $OUT 2>>gpurun_out/r02_ar_micro_err.log
python tools/r02/ar_micro.py --tag world1 --world 1 --sizes 4096,4377904,67108864 --algos oneshot >> $OUT 2>>gpurun_out/r02_ar_micro_err.log
cat $OUT
tail -5 gpurun_out/r02_ar_micro_err.log
