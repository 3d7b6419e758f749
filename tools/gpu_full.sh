#!/bin/bash
# Full round check: smoke, all GPU tests, default bench (with cpu baseline), fp32 bench, reference arm, ncu launch list.
# Usage: gpurun --timeout 1500 -- bash tools/gpu_full.sh <tag>
TAG=${1:-full}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -3
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tee $OUT/pytest_gpu.log | tail -5
echo "== bench"; timeout 600 python bench.py 2>&1 | tee $OUT/bench.log | tail -1 | cut -c1-1500
echo "== bench fp32"; timeout 300 python bench.py --precision float --no-check 2>&1 | tee $OUT/bench_f32.log | tail -1 | cut -c1-600
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee $OUT/bench_ref.log | tail -1 | cut -c1-600
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-check > $OUT/ncu_launches.log 2>&1
ls -la $OUT
