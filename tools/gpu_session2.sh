#!/bin/bash
# Usage: gpurun --timeout 1800 -- bash tools/gpu_session2.sh <tag> [pytest-filter]
TAG=${1:-r01b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $OUT/smoke.log | tail -3
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} 2>&1 | tee $OUT/pytest_gpu.log | tail -12
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tee $OUT/bench.log | tail -2
echo "== bench psi table"; timeout 600 python bench.py --steps 10 --warmup 3 --psi-table --no-check 2>&1 | tee $OUT/bench_table.log | tail -2
echo "== bench fp32"; timeout 600 python bench.py --steps 10 --warmup 3 --precision float --no-check 2>&1 | tee $OUT/bench_f32.log | tail -2
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-check > $OUT/ncu_launches.log 2>&1
echo "== ncu full (tile kernels, M=2e6)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spread_tile|interp_tile" -c 2 \
    -o $OUT/prof_tile python bench.py --steps 1 --warmup 3 --no-check --nodes 2000000 > $OUT/ncu_full.log 2>&1
ls -la $OUT
