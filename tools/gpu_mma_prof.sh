#!/bin/bash
# Usage: gpurun --timeout 900 -- bash tools/gpu_mma_prof.sh <tag> [extra bench args for the ncu pass]
TAG=${1:-mma}; shift
bash tools/gpu_mma.sh $TAG
OUT=gpurun_out/$TAG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"interp_mma_kernel|spread_mma_kernel" -s 6 -c 2 \
    -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-check "$@" > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-200
