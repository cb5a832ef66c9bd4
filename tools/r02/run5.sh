#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batcher_gpu.py -x -q > gpurun_out/r02_pytest_hpb.log 2>&1
tail -5 gpurun_out/r02_pytest_hpb.log
OUT=gpurun_out/r02_gather_micro.jsonl
: > $OUT
python tools/r02/gather_micro.py --tag default >> $OUT 2>gpurun_out/r02_gather_err.log
MB_TMA_TABLE_CONTIG=0 python tools/r02/gather_micro.py --tag strided >> $OUT 2>>gpurun_out/r02_gather_err.log
MB_COPY_IMPL=ldg python tools/r02/gather_micro.py --tag ldg >> $OUT 2>>gpurun_out/r02_gather_err.log
python tools/r02/gather_micro.py --tag e1024 --envs 1024 --reps 6 >> $OUT 2>>gpurun_out/r02_gather_err.log
python tools/r02/gather_micro.py --tag e64 --envs 64 >> $OUT 2>>gpurun_out/r02_gather_err.log
python tools/r02/gather_micro.py --tag e4096 --envs 4096 --reps 5 >> $OUT 2>>gpurun_out/r02_gather_err.log
cat $OUT; tail -3 gpurun_out/r02_gather_err.log
