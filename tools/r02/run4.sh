#!/usr/bin/env bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_gather_micro.jsonl
: > $OUT
python tools/r02/gather_micro.py --tag default >> $OUT 2>gpurun_out/r02_gather_err.log
MB_TMA_TABLE_CONTIG=0 python tools/r02/gather_micro.py --tag strided >> $OUT 2>>gpurun_out/r02_gather_err.log
MB_COPY_IMPL=ldg python tools/r02/gather_micro.py --tag ldg >> $OUT 2>>gpurun_out/r02_gather_err.log
MB_TMA_STAGES=6 MB_TMA_TILE=8192 MB_TMA_WARPS=4 python tools/r02/gather_micro.py --tag w4s6t8k >> $OUT 2>>gpurun_out/r02_gather_err.log
MB_TMA_WARPS=2 MB_TMA_STAGES=6 python tools/r02/gather_micro.py --tag w2s6 >> $OUT 2>>gpurun_out/r02_gather_err.log
MB_TMA_INLINE_CONTIG=1 python tools/r02/gather_micro.py --tag inline_contig >> $OUT 2>>gpurun_out/r02_gather_err.log
python tools/r02/gather_micro.py --tag e1024 --envs 1024 --reps 6 >> $OUT 2>>gpurun_out/r02_gather_err.log
cat $OUT; tail -3 gpurun_out/r02_gather_err.log
