#!/bin/bash
# Usage: gpurun --timeout 900 -- bash tools/gpu_quick2.sh <tag>   (kernel iteration: 3-D parity subset + bench lines)
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "tile3d or cfg3 or images or live or golden or adjointness or ndft_fixture" 2>&1 | tee $OUT/pytest.log | tail -4
timeout 300 python bench.py --steps 10 2>&1 | tee $OUT/bench.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('fp64', d['ms_per_step'], d['kernel_ms'], d['stage_ms']['adjoint'], d['rel_l2']['trafo'], d['rel_l2']['adjoint'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"
timeout 300 python bench.py --steps 10 --precision float 2>&1 | tee $OUT/bench_f32.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('fp32', d['ms_per_step'], d['kernel_ms'], d['stage_ms']['adjoint'], d['rel_l2']['trafo'], d['rel_l2']['adjoint'], 'e2e', d['e2e']['ms_per_step'])"
