#!/bin/bash
# Usage: gpurun -- bash tools/gpu_launches.sh <tag> [bench args]
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-check "$@" > $OUT/ncu_launches.log 2>&1
tail -2 $OUT/ncu_launches.log | cut -c1-200
