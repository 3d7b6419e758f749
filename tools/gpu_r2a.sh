#!/bin/bash
# Usage: gpurun --timeout 2400 -- bash tools/gpu_r2a.sh <tag>
# Round-2 session A: new full-size / multi-GPU parity tests, the whole GPU suite, bench lines with rel_l2 and the
# measured reference arm, ncu launch list + full captures of the final fp64 and fp32 tensor kernels.
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt; nproc >> $OUT/gpu.txt; free -g >> $OUT/gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -3
echo "== pytest new"; timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fullsize.py -m gpu -x -q -s --durations=10 2>&1 | tee $OUT/pytest_new.log | tail -25
echo "== bench"; timeout 900 python bench.py 2>&1 | tee $OUT/bench.log | tail -1 | cut -c1-1500
echo "== bench fp32"; timeout 600 python bench.py --precision float 2>&1 | tee $OUT/bench_f32.log | tail -1 | cut -c1-600
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tee $OUT/bench_ref.log | tail -1 | cut -c1-600
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tee $OUT/pytest_gpu.log | tail -5
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-check > $OUT/ncu_launches.log 2>&1
echo "== ncu full fp64"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"interp_mma_kernel|spread_mma_kernel" -s 6 -c 2 \
    -o $OUT/prof_fp64 python bench.py --steps 1 --warmup 3 --no-check > $OUT/ncu_full64.log 2>&1
echo "== ncu full fp32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"interp_tf32_kernel|spread_tf32_kernel" -s 6 -c 2 \
    -o $OUT/prof_fp32 python bench.py --precision float --steps 1 --warmup 3 --no-check > $OUT/ncu_full32.log 2>&1
ls -la $OUT
