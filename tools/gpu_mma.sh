#!/bin/bash
# Usage: gpurun --timeout 600 -- bash tools/gpu_mma.sh <tag>
TAG=${1:-mma}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -x -q -k "tile3d or cfg3_grid" 2>&1 | tee $OUT/pytest.log | tail -15
summ() { python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); s=d['stage_ms']; print('$1 value %.3e ms %.2f B %.2f BT %.2f e2e %.1f ms'%(d['value'],d['ms_per_step'],s['trafo']['B'],s['adjoint']['BT'],d['e2e']['ms_per_step']))
    else: print(ln.rstrip()[:300])"; }
timeout 90 python bench.py --steps 10 --warmup 3 --no-check 2>&1 | tee $OUT/bench.log | summ mma
timeout 90 python bench.py --steps 10 --warmup 3 --no-check --b-flush 2 2>&1 | tee $OUT/bench_tma.log | summ mma-tma
