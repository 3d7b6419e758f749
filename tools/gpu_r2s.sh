#!/bin/bash
# Usage: gpurun --timeout 500 -- bash tools/gpu_r2s.sh <tag>
# compute-sanitizer memcheck over small cases of the tcgen05 kernels (tc5.cu), and the fp32 bench line with the final labels.
TAG=${1:-r2s}; OUT=gpurun_out/$TAG; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 150 $CS --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_features.py -m gpu -q -x -p no:cacheprovider \
    -k "tc5 and (n32_m6 or nonpow2_m4 or m5_sparse)" > $OUT/mem_tc5.log 2>&1
echo "mem_tc5: rc=$? $(grep 'ERROR SUMMARY' $OUT/mem_tc5.log | tail -1) | $(tail -1 $OUT/mem_tc5.log)"
echo "== bench fp32"; timeout 120 python bench.py --steps 10 --warmup 3 --precision float 2>$OUT/bench_f32.err | tee $OUT/bench_fp32.json | cut -c1-200
ls -la $OUT
