#!/bin/bash
# Usage: gpurun -- bash tools/gpu_variants.sh <tag> v1 v2 ...   (libs prebuilt under nfft_b200/lib_var/<v>/libnfftcu.so)
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
cp nfft_b200/lib/libnfftcu.so /tmp/libnfftcu_base.so
for v in base "$@"; do
  if [ $v = base ]; then cp /tmp/libnfftcu_base.so nfft_b200/lib/libnfftcu.so; else cp nfft_b200/lib_var/$v/libnfftcu.so nfft_b200/lib/libnfftcu.so; fi
  timeout 120 python bench.py --steps 10 --warmup 3 --no-check $BENCH_ARGS 2>&1 | tee $OUT/bench_$v.log | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); s=d['stage_ms']; k=d['kernel_ms']; print('$v: ms %.2f B %.3f BT %.3f (stage B %.2f BT %.2f)'%(d['ms_per_step'],k['B'],k['BT'],s['trafo']['B'],s['adjoint']['BT']))
"
done
cp /tmp/libnfftcu_base.so nfft_b200/lib/libnfftcu.so
