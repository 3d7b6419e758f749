#!/bin/bash
# Usage: gpurun --timeout 1200 -- bash tools/gpu_prof.sh <tag> <kernel-regex> [bench args...]
TAG=$1; REGEX=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1000 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s ${SKIP:-6} -c ${COUNT:-2} \
    -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-check "$@" > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT
