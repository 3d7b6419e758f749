#!/bin/bash
# Usage: gpurun --timeout 1500 -- bash tools/gpu_sanitize.sh <tag>
# compute-sanitizer over small cases of the kernels added in round 2 (memcheck everywhere, racecheck on the
# shared-memory FFT / tile kernels); summaries under gpurun_out/<tag>/.
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool pytest-args...
  local name=$1 tool=$2; shift 2
  timeout 600 $CS --tool $tool --error-exitcode 7 --print-limit 5 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > $OUT/$name.log 2>&1
  echo "$name ($tool): rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' $OUT/$name.log) clean-summaries; $(grep 'ERROR SUMMARY' $OUT/$name.log | tail -1) | $(tail -1 $OUT/$name.log)"
}
run mem_features memcheck tests/test_gpu_features.py -k "(batched_transforms and (2d_tiles-3 or 1d_generic-3 or 2d_nonpow2-3) and double) or (device_mri and 32-12) or (adjoint_mul and 2-N0) or (gaussian and 2d_fg-double) or (batched_solver and CGNR-double) or plan_cache"
run mem_multi memcheck tests/test_gpu_multi.py -k "(group_vs_oracle and 1-) or fused_reduce"
run mem_fft memcheck tests/test_gpu_parity.py -k "fft_general and double and (N2- or N3- or N4- or N5- or N6- or N7-)"
run race_fft racecheck tests/test_gpu_parity.py -k "fft_general and double and (N2- or N6- or N7-)"
run race_batch racecheck tests/test_gpu_features.py -k "batched_transforms and 2d_tiles-3 and double"
ls -la $OUT
