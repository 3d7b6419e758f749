#!/bin/bash
# Usage: gpurun --timeout 900 -- bash tools/gpu_prof_new.sh <tag>
TAG=${1:-pn}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 800 ncu --set full --clock-control none --import-source on \
  -k regex:"fft_bluestein_kernel|deconv_crop_reduce_kernel|slab_scatter_kernel|slab_gather_kernel|peer_barrier_kernel|interp_tile2_kernel|spread_tile2_kernel" \
  -c 40 -o $OUT/prof_new python tools/prof_new_kernels.py > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log; ls -la $OUT
