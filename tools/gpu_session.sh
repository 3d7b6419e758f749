#!/bin/bash
# One GPU-box session: smoke, GPU tests, bench, ncu launch list + one full capture, microbench.
# Usage: gpurun --timeout 1500 -- bash tools/gpu_session.sh [tag]
TAG=${1:-r01a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -5
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tee $OUT/pytest_gpu.log | tail -15
echo "== microbench"; timeout 120 ./tools/microbench 2>&1 | tee $OUT/microbench.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee $OUT/bench.log | tail -3
echo "== bench fp32"; timeout 600 python bench.py --steps 5 --warmup 3 --precision float --no-check 2>&1 | tee $OUT/bench_f32.log | tail -2
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-check > $OUT/ncu_launches.log 2>&1
echo "== ncu full (spread+interp, M=1e6)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spread_generic|interp_generic" -c 2 \
    -o $OUT/prof_B python bench.py --steps 1 --warmup 3 --no-check --nodes 1000000 > $OUT/ncu_full.log 2>&1
ls -la $OUT
