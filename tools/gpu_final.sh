#!/bin/bash
# Usage: gpurun --timeout 2400 -- bash tools/gpu_final.sh <tag>
# Round-end session: smoke, the whole GPU suite, bench lines (fp64 with cpu_baseline, fp32, reference arm), the ncu
# launch list of the bench command and one ncu --set full capture of the dominant kernels.
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -3
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tee $OUT/pytest_gpu.log | tail -5
echo "== bench"; timeout 600 python bench.py 2>&1 | tee $OUT/bench.log | tail -1 | cut -c1-200
echo "== bench fp32"; timeout 300 python bench.py --precision float --no-check 2>&1 | tee $OUT/bench_f32.log | tail -1 | cut -c1-200
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee $OUT/bench_ref.log | tail -1 | cut -c1-200
echo "== configs"; timeout 400 python tools/bench_configs.py --configs cfg1,cfg2,cfg3r,cfg5 2>&1 | tee $OUT/configs.log | cut -c1-160
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-check > $OUT/ncu_launches.log 2>&1
echo "== ncu full (B / B^T tensor kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"interp_mma_kernel|spread_mma_kernel" -s 6 -c 2 \
    -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-check > $OUT/ncu_full.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt
ls -la $OUT
