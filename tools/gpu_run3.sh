#!/usr/bin/env bash
# copy kernel tuning pass (1 GPU)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_copy_gpu.py -m gpu -x -q 2>&1 | tail -3
rm -f gpurun_out/sweep_copy.json
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 600 python tools/sweep_copy.py --envs 256 1024 4096 --reps 10 --tag $tag 2>&1 | grep -E '"op": "(stack|cat_1launch|contig)"' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['impl'], r['op'], r['envs'], r['gbs'], r['best_gbs'], r['frac_hbm'])"
}
run ldg8 MB_COPY_IMPL=ldg
run ldg16 MB_COPY_IMPL=ldg MB_COPY_CTAS_PER_SM=16
run tma_8k_w4_s6 MB_COPY_IMPL=tma
run tma_16k_w4_s3 MB_COPY_IMPL=tma MB_TMA_TILE=16384 MB_TMA_STAGES=3 MB_TMA_STORES=1
run tma_16k_w3_s4 MB_COPY_IMPL=tma MB_TMA_TILE=16384 MB_TMA_WARPS=3 MB_TMA_STAGES=4 MB_TMA_STORES=2
run tma_8k_w6_s4 MB_COPY_IMPL=tma MB_TMA_WARPS=6 MB_TMA_STAGES=4 MB_TMA_STORES=2
run tma_4k_w6_s8 MB_COPY_IMPL=tma MB_TMA_TILE=4096 MB_TMA_WARPS=6 MB_TMA_STAGES=8 MB_TMA_STORES=4
run tma_8k_w4_s6_st2 MB_COPY_IMPL=tma MB_TMA_STORES=2
run tma_32k_w2_s3 MB_COPY_IMPL=tma MB_TMA_TILE=32768 MB_TMA_WARPS=2 MB_TMA_STAGES=3 MB_TMA_STORES=1
run auto
