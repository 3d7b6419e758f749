#!/bin/bash
# Usage: gpurun --timeout 600 -- bash tools/gpu_r2o.sh <tag>
# Round-2 closing session on one GPU: the tc5 test, fp64 + fp32 bench lines, ncu capture of the tcgen05 interpolation kernel
# and the fp32 launch list.
TAG=${1:-r2o}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_features.py -m gpu -q -k "tc5 or slab" --timeout 120 2>&1 | tail -3 | tee $OUT/pytest_tc5.log
echo "== bench fp64"; timeout 300 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench_fp64.json | cut -c1-400
echo "== bench fp32"; timeout 200 python bench.py --steps 10 --warmup 3 --precision float 2>$OUT/bench_f32.err | tee $OUT/bench_fp32.json | cut -c1-400
echo "== ncu full tc5"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"tc5_interp_kernel|spread_tf32_kernel" -c 2 \
    -o $OUT/prof_tc5 python tools/tc5_check.py --big > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
echo "== ncu launches fp32"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_fp32.csv \
    python bench.py --steps 2 --warmup 3 --no-check --precision float > $OUT/ncu_launches.log 2>&1
ls -la $OUT
