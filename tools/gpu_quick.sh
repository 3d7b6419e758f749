#!/bin/bash
# Usage: gpurun --timeout 900 -- bash tools/gpu_quick.sh <tag> [pytest -k filter]
TAG=${1:-quick}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 150 python -m pytest tests -m gpu -x -q -k "${2:-tile3d or test_stages or oracle or golden}" 2>&1 | tee $OUT/pytest.log | tail -4
timeout 90 python bench.py --steps 10 --warmup 3 --no-check 2>&1 | tee $OUT/bench.log | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); s=d['stage_ms']; print('fp64 value %.3e ms %.2f B %.2f BT %.2f e2e %.1f ms'%(d['value'],d['ms_per_step'],s['trafo']['B'],s['adjoint']['BT'],d['e2e']['ms_per_step']))
    else: print(ln.rstrip()[:300])"
timeout 90 python bench.py --steps 10 --warmup 3 --no-check --precision float 2>&1 | tee $OUT/bench_f32.log | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); s=d['stage_ms']; print('fp32 value %.3e ms %.2f B %.2f BT %.2f'%(d['value'],d['ms_per_step'],s['trafo']['B'],s['adjoint']['BT']))
    else: print(ln.rstrip()[:300])"
