#!/bin/bash
# Usage: gpurun --timeout 600 -- bash tools/gpu_r2r.sh <tag>
# Closing session of round 2 on one GPU: tc5 tests, smoke, fp64 + fp32 bench lines, ncu capture of both tcgen05 kernels.
TAG=${1:-r2r}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_features.py -m gpu -q -k "tc5" --timeout 100 2>&1 | tail -3 | tee $OUT/pytest_tc5.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
echo "== bench fp64"; timeout 200 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench_fp64.json | cut -c1-200
echo "== bench fp32"; timeout 200 python bench.py --steps 10 --warmup 3 --precision float 2>$OUT/bench_f32.err | tee $OUT/bench_fp32.json | cut -c1-200
echo "== ncu full tc5"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"tc5_interp_kernel|tc5_spread_kernel" -c 2 \
    -o $OUT/prof_tc5 python tools/tc5_check.py --big > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
ls -la $OUT
