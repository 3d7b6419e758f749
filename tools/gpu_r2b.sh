#!/bin/bash
# Usage: gpurun --timeout 2400 -- bash tools/gpu_r2b.sh <tag>
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -3
echo "== pytest new"; timeout 1800 python -m pytest tests/test_gpu_features.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -m gpu -q -s --durations=8 2>&1 | tee $OUT/pytest_new.log | tail -30
echo "== pytest gpu parity"; timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tee $OUT/pytest_parity.log | tail -8
echo "== bench"; timeout 900 python bench.py 2>&1 | tee $OUT/bench.log | tail -1 | cut -c1-300
echo "== bench fp32"; timeout 600 python bench.py --precision float 2>&1 | tee $OUT/bench_f32.log | tail -1 | cut -c1-300
ls -la $OUT
